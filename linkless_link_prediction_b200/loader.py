"""Index batching with the reference's RNG stream.

The reference iterates ``DataLoader(range(n), batch_size, shuffle=True)`` (train_teacher_gnn.py:35,
main.py:72-73,167-168).  Collating 65,536 Python ints per batch costs more host time than the whole GPU step,
so ``shuffled_batches`` reproduces the exact index stream without the Python lists.  The draws from the global
torch generator happen at the same moments as in the DataLoader: the base seed when the iterator is created
(``iter(DataLoader(...))``), the permutation seed of ``RandomSampler`` at the first ``next()``, then
``torch.randperm(n, generator=g)``.  ``tests/test_host_logic.py`` pins the equality against the real DataLoader.
"""
from __future__ import annotations

import torch


class shuffled_batches:
    """Iterator equal to ``iter(DataLoader(range(n), batch_size, shuffle=True))`` (CPU int64 index tensors)."""

    def __init__(self, n: int, batch_size: int):
        self.n, self.batch_size = int(n), int(batch_size)
        torch.empty((), dtype=torch.int64).random_()  # _BaseDataLoaderIter._base_seed, drawn at iter() time
        self._perm = None
        self._pos = 0

    def __iter__(self):
        return self

    def __next__(self) -> torch.Tensor:
        if self._perm is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())  # RandomSampler.__iter__
            g = torch.Generator()
            g.manual_seed(seed)
            self._perm = torch.randperm(self.n, generator=g)
        if self._pos >= self.n:
            raise StopIteration
        out = self._perm[self._pos:self._pos + self.batch_size]
        self._pos += self.batch_size
        return out


def sequential_batches(n: int, batch_size: int):
    """``DataLoader(range(n), batch_size)`` without shuffling (the scoring loops, train_teacher_gnn.py:95)."""
    idx = torch.arange(n)
    for s in range(0, n, batch_size):
        yield idx[s:s + batch_size]
