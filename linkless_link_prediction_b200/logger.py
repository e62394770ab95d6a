"""Result bookkeeping with the reference's semantics (src/logger.py:3-89): per-run lists of
(valid, test[, old_old, old_new, new_new]) tuples and best-by-validation statistics.  Pure host code."""
import torch


class Logger(object):
    width = 2

    def __init__(self, runs, info=None):
        self.info = info
        self.results = [[] for _ in range(runs)]

    def add_result(self, run, result):
        assert len(result) == self.width
        assert 0 <= run < len(self.results)
        self.results[run].append(result)

    def reset(self, run):
        assert 0 <= run < len(self.results)
        self.results[run] = []

    def best(self, run):
        r = 100 * torch.tensor(self.results[run])
        arg = r[:, 0].argmax().item()
        return r, arg

    def print_statistics(self, run=None):
        if run is not None:
            result = torch.tensor(self.results[run])
            argmax = result[:, 0].argmax().item()
            print(f'Run {run + 1:02d}:')
            print(f'Highest Valid: {result[:, 0].max():.4f}')
            print(f'   Final Test: {result[argmax, 1]:.4f}')
        else:
            rows = []
            for i in range(len(self.results)):
                r, arg = self.best(i)
                rows.append((r[:, 0].max().item(), r[arg, 1].item()))
            best = torch.tensor(rows)
            print('All runs:')
            print(f'Highest Valid: {best[:, 0].mean():.2f} ± {best[:, 0].std():.2f}')
            print(f'   Final Test: {best[:, 1].mean():.2f} ± {best[:, 1].std():.2f}')


class ProductionLogger(Logger):
    width = 5
    names = ('Final val', 'Final Test', 'old_old Test', 'old_new Test', 'new_new Test')

    def print_statistics(self, run=None):
        if run is not None:
            r, arg = self.best(run)
            print(f'Run {run + 1:02d}:')
            for j, name in enumerate(self.names):
                print(f'  {name}: {r[arg, j]:.2f}')
        else:
            rows = []
            for i in range(len(self.results)):
                r, arg = self.best(i)
                rows.append(tuple(r[arg, j].item() for j in range(5)))
            best = torch.tensor(rows)
            print('All runs:')
            for j, name in enumerate(self.names):
                print(f'  {name}: {best[:, j].mean():.2f} ± {best[:, j].std():.2f}')
