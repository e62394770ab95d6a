"""Generate the golden fixtures by running the REFERENCE'S OWN functions in this container.

Run here (needs ``/root/reference``; it is absent on the GPU box, which only reads the
committed ``*.pt`` fixtures):

    python tests/golden/make_golden.py

How: the reference imports torch_geometric / torch_sparse / torch_cluster / ogb at module top
(``src/train_teacher_gnn.py:7-16``, ``src/main.py:9-22``, ``src/models.py:3``), none of which
exist here.  We register stub modules for them whose hot-path entry points are the oracle's
restatements (``SAGEConv``, ``negative_sampling``, ``random_walk``, ``Evaluator``); everything
else the reference executes — ``models.py`` (MLP / SAGE / LinkPredictor), ``train()``,
``test_transductive()``, ``kl_loss``, the LLP_D/LLP_R block and the optimiser tail of
``main.train()`` — is the reference's unmodified code.  ``main.py`` runs ``main()`` at import
(``:515``) and hard-codes ``"cuda"`` (``:50,:191``), so its functions are extracted with ``ast``
and ``Tensor.to("cuda")`` is mapped to a no-op while they run.
"""
import argparse
import ast
import os
import random
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
sys.path.insert(0, ROOT)

from oracle import llp_oracle as O  # noqa: E402


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Anything()


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


def _install_stubs():
    names = [
        "torch_geometric", "torch_geometric.utils", "torch_geometric.transforms", "torch_geometric.nn",
        "torch_geometric.nn.conv", "torch_geometric.nn.dense", "torch_geometric.nn.dense.linear",
        "torch_geometric.typing", "torch_geometric.data", "torch_geometric.datasets", "torch_geometric.seed",
        "torch_sparse", "torch_cluster", "ogb", "ogb.linkproppred",
    ]
    mods = {n: _Stub(n) for n in names}
    for n, m in mods.items():
        sys.modules[n] = m
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(mods[parent], child, m)
    mods["torch_geometric.nn"].SAGEConv = O.SAGEConv
    mods["torch_geometric.nn.conv"].MessagePassing = torch.nn.Module
    mods["torch_geometric.nn.dense.linear"].Linear = torch.nn.Linear

    def negative_sampling(edge_index, num_nodes=None, num_neg_samples=None, method="sparse"):
        assert method == "dense"
        return O.negative_sampling_dense(edge_index, num_nodes, num_neg_samples)

    mods["torch_geometric.utils"].negative_sampling = negative_sampling
    mods["torch_cluster"].random_walk = O.random_walk
    mods["ogb.linkproppred"].Evaluator = O.Evaluator

    def seed_everything(seed):
        random.seed(seed)
        import numpy as np
        np.random.seed(seed)
        torch.manual_seed(seed)

    mods["torch_geometric.seed"].seed_everything = seed_everything
    mods["torch_geometric"].seed = mods["torch_geometric.seed"]
    return seed_everything


def _load_main_functions():
    """Functions of ``src/main.py`` without executing its module-level ``main()`` call."""
    src = open(os.path.join(REF, "main.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom, ast.FunctionDef)) and
            getattr(n, "name", "") != "main"]
    mod = types.ModuleType("ref_main")
    exec(compile(ast.Module(body=keep, type_ignores=[]), os.path.join(REF, "main.py"), "exec"), mod.__dict__)
    return mod


class _CudaIsCpu:
    """``x.to("cuda")`` -> ``x`` while the reference's student step runs on this CPU-only box."""

    def __enter__(self):
        self._orig = torch.Tensor.to

        def to(t, *a, **k):
            if a and isinstance(a[0], str) and a[0].startswith("cuda"):
                return t
            return self._orig(t, *a, **k)

        torch.Tensor.to = to

    def __exit__(self, *exc):
        torch.Tensor.to = self._orig


def tiny_graph(n=240, und=900, f=40, seed=0):
    ei = O.synthetic_undirected_graph(n, und, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = (torch.rand(n, f, generator=g) < 0.15).float() * torch.rand(n, f, generator=g)
    # hold out ~15% of the undirected pairs as valid/test positives; equal-sized random negatives
    und_mask = ei[0] < ei[1]
    pairs = ei[:, und_mask].t()
    perm = torch.randperm(pairs.size(0), generator=g)
    n_val, n_test = 40, 80
    val, test, train = pairs[perm[:n_val]], pairs[perm[n_val:n_val + n_test]], pairs[perm[n_val + n_test:]]
    train_dir = torch.cat([train, train.flip(1)], 0)
    order = torch.argsort(train_dir[:, 0] * n + train_dir[:, 1])
    train_dir = train_dir[order]
    split = {
        "train": {"edge": train_dir},
        "valid": {"edge": val, "edge_neg": torch.randint(0, n, (n_val, 2), generator=g)},
        "test": {"edge": test, "edge_neg": torch.randint(0, n, (n_test, 2), generator=g)},
    }
    return x, split


def clone_sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def main():
    seed_everything = _install_stubs()
    sys.path.insert(0, REF)
    import models as ref_models  # the reference's models.py
    import train_teacher_gnn as ref_teacher  # the reference's teacher script (main() is guarded)
    ref_main = _load_main_functions()
    out = {}

    # ---- G1: models.py forward known-answers (eval mode) --------------------------------
    torch.manual_seed(11)
    x = torch.randn(50, 24)
    ei = O.synthetic_undirected_graph(50, 120, seed=3)
    mlp = ref_models.MLP(3, 24, 32, 16, 0.5).eval()
    sage = ref_models.SAGE("cora", 24, 32, 16, 3, 0.5, O.SAGEConv).eval()
    sage_u = ref_models.SAGE("coauthor-physics", 24, 32, 16, 2, 0.5, O.SAGEConvUpdated).eval()
    pred = ref_models.LinkPredictor("mlp", 16, 32, 1, 3, 0.5).eval()
    pred_in = ref_models.LinkPredictor("inner", 16, 32, 1, 2, 0.5).eval()
    xi, xj = torch.randn(70, 16), torch.randn(70, 16)
    x3i, x3j = torch.randn(9, 5, 16), torch.randn(9, 5, 16)
    with torch.no_grad():
        out["models"] = {
            "x": x, "edge_index": ei, "xi": xi, "xj": xj, "x3i": x3i, "x3j": x3j,
            "mlp_sd": clone_sd(mlp), "mlp_out": mlp(x),
            "sage_sd": clone_sd(sage), "sage_out": sage(x, ei),
            "sage_u_sd": clone_sd(sage_u), "sage_u_out": sage_u(x, ei),
            "pred_sd": clone_sd(pred), "pred_out": pred(xi, xj), "pred_out3": pred(x3i, x3j),
            "pred_in_sd": clone_sd(pred_in), "pred_in_out": pred_in(xi, xj),
        }

    # ---- G2: kl_loss / cosine_loss (main.py:24-31) --------------------------------------
    torch.manual_seed(12)
    s, t = torch.rand(64, 12), torch.rand(64, 12)
    out["kl"] = {"s": s, "t": t, "kl_T1": ref_main.kl_loss(s, t, 1), "kl_T2": ref_main.kl_loss(s, t, 2.0),
                 "cos": ref_main.cosine_loss(s, t)}

    # ---- G3: teacher epoch + eval (train_teacher_gnn.py:21-155), dropout 0 ---------------
    x, split = tiny_graph()
    H = 32

    class D:  # minimal Data-like container
        pass

    data = D()
    data.x = x
    data.adj_t = split["train"]["edge"].t().contiguous()
    for batch_size, tag in ((64 * 1024, "teacher_fullbatch"), (512, "teacher_minibatch")):
        seed_everything(5)
        model = ref_models.SAGE("cora", x.size(1), H, H, 2, 0.0, O.SAGEConv)
        predictor = ref_models.LinkPredictor("mlp", H, H, 1, 2, 0.0)
        sd0 = {"gnn": clone_sd(model), "predictor": clone_sd(predictor)}
        opt = torch.optim.Adam(list(model.parameters()) + list(predictor.parameters()), lr=0.005)
        seed_everything(6)
        losses = [ref_teacher.train(model, predictor, data, split, opt, batch_size, "sage", "cora", "transductive")
                  for _ in range(2)]
        args = argparse.Namespace(minibatch=False)
        results, h = ref_teacher.test_transductive(model, predictor, data, split, O.Evaluator("ogbl-ddi"),
                                                   batch_size, "sage", "cora", args)
        out[tag] = {"x": x, "split": split, "H": H, "batch_size": batch_size, "sd0": sd0, "seed_init": 5,
                    "seed_train": 6, "lr": 0.005, "losses": losses, "results": results, "h": h,
                    "sd1": {"gnn": clone_sd(model), "predictor": clone_sd(predictor)}}
        teacher_model, teacher_pred = model, predictor

    # ---- G4: student KD step (main.py:147-236): LLP_D + LLP_R + True_label ---------------
    teacher_model.eval()
    with torch.no_grad():
        t_h = teacher_model(data.x, data.adj_t)
    for p in teacher_pred.parameters():
        p.requires_grad = False
    args = argparse.Namespace(transductive="transductive", node_batch_size=120, link_batch_size=800,
                              LLP_R=1.0, LLP_D=1.0, True_label=1.0, KD_RM=0.0, KD_LM=0.0, margin=0.1,
                              rw_step=3, ps_method="nb", ns_rate=1, hops=2, datasets="cora")
    seed_everything(7)
    student = ref_models.MLP(2, x.size(1), H, H, 0.0)
    s_pred = ref_models.LinkPredictor("mlp", H, H, 1, 2, 0.0)
    sd0 = {"mlp": clone_sd(student), "predictor": clone_sd(s_pred)}
    opt = torch.optim.Adam(list(student.parameters()) + list(s_pred.parameters()), lr=0.01)
    seed_everything(8)
    with _CudaIsCpu():
        losses = [ref_main.train(student, s_pred, t_h, teacher_pred, data, split, opt, args, torch.device("cpu"))
                  for _ in range(2)]
    out["student"] = {"x": x, "split": split, "H": H, "t_h": t_h, "teacher_pred_sd": clone_sd(teacher_pred),
                      "args": vars(args), "sd0": sd0, "seed_init": 7, "seed_train": 8, "lr": 0.01,
                      "losses": losses, "sd1": {"mlp": clone_sd(student), "predictor": clone_sd(s_pred)}}

    # ---- G5: the LLP_R block alone, lifted out of main.train with its own variable names ---
    torch.manual_seed(13)
    t_r, s_r = torch.rand(33, 12, 1), torch.rand(33, 12, 1)
    import itertools
    import numpy as np
    margin = 0.1
    dim_pairs = np.array([p for p in itertools.combinations(range(12), r=2)]).T
    teacher_rank_list = torch.zeros((len(t_r), dim_pairs.shape[1], 1))
    teacher_rank_list[t_r[:, dim_pairs[0]] > (t_r[:, dim_pairs[1]] + margin)] = 1
    teacher_rank_list[t_r[:, dim_pairs[0]] < (t_r[:, dim_pairs[1]] - margin)] = -1
    llp_r = torch.nn.MarginRankingLoss(margin=margin)(s_r[:, dim_pairs[0]].squeeze(), s_r[:, dim_pairs[1]].squeeze(),
                                                     teacher_rank_list.squeeze())
    out["llp_r"] = {"s_r": s_r.squeeze(-1), "t_r": t_r.squeeze(-1), "margin": margin, "loss": llp_r}

    torch.save(out, os.path.join(HERE, "reference_golden.pt"))
    print("wrote", os.path.join(HERE, "reference_golden.pt"), os.path.getsize(os.path.join(HERE, "reference_golden.pt")), "bytes")
    print({k: (v["losses"] if isinstance(v, dict) and "losses" in v else "") for k, v in out.items()})


if __name__ == "__main__":
    main()
