#!/usr/bin/env python
"""fp32-mode accuracy A/B at config sizes: one training step (forward, BCE, backward) of the C5 hub slice and of the C4
collab shape through (a) the 3xTF32 tensor-core GEMMs (default) and (b) the CUDA-core fp32 GEMMs (llp_set_tuning(21, 1)),
each against the fp64 oracle, next to the fp32 oracle's own distance from fp64.  Development / evidence tool:
    python tools/fp32_accuracy.py [c5] [c4] > profiles/r02_fp32_accuracy.txt"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import linkless_link_prediction_b200 as L  # noqa: E402
from linkless_link_prediction_b200 import _native as N  # noqa: E402
from linkless_link_prediction_b200 import ops  # noqa: E402
from linkless_link_prediction_b200.data import synthetic_dataset  # noqa: E402
from oracle import llp_oracle as O  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def run(name, x, adj, pos, neg, f, H, layers):
    torch.manual_seed(0)
    mo = O.SAGE("cfg", f, H, H, layers, 0.0, O.SAGEConv)
    po = O.LinkPredictor("mlp", H, H, 1, 2, 0.0)
    edges = torch.cat((pos, neg), dim=-1)
    label = torch.cat((torch.ones(pos.size(1)), torch.zeros(neg.size(1))))

    def oracle(m, p, xx, lab):
        for q in list(m.parameters()) + list(p.parameters()):
            q.grad = None
        h = m(xx, adj)
        loss = O.bce_loss(p(h[edges[0]], h[edges[1]]).squeeze(), lab)
        loss.backward()
        return h.detach(), float(loss), [q.grad.clone() for q in list(m.parameters()) + list(p.parameters())]

    h64, l64, g64 = oracle(copy.deepcopy(mo).double(), copy.deepcopy(po).double(), x.double(), label.double())
    h32, l32, g32 = oracle(mo, po, x, label)
    names = [k for k, _ in list(mo.named_parameters())] + ["pred." + k for k, _ in po.named_parameters()]
    ops.set_compute_dtype(torch.float32)
    res = {}
    for tag, knob in (("tf32x3", 0), ("simt", 1)):
        N.load().llp_set_tuning(21, knob)
        md = L.SAGE("cfg", f, H, H, layers, 0.0, L.SAGEConv); md.load_state_dict(mo.state_dict()); md.to(dev).train()
        pd = L.LinkPredictor("mlp", H, H, 1, 2, 0.0); pd.load_state_dict(po.state_dict()); pd.to(dev).train()
        hd = md(x.to(dev), adj.to(dev))
        ed = edges.to(dev)
        ld = ops.bce_loss(pd.score(hd, ed[0].contiguous(), ed[1].contiguous()).reshape(-1), pos.size(1))
        ld.backward()
        res[tag] = (hd.detach().float().cpu(), float(ld.detach()), [q.grad.float().cpu() for q in list(md.parameters()) + list(pd.parameters())])
    N.load().llp_set_tuning(21, 0)
    print(f"== {name}: relative error (norm) against the fp64 oracle")
    print(f"{'':28s} {'oracle fp32':>12s} {'cuda tf32x3':>12s} {'cuda simt':>12s}")
    print(f"{'loss':28s} {abs(l32 - l64) / abs(l64):12.2e} {abs(res['tf32x3'][1] - l64) / abs(l64):12.2e} {abs(res['simt'][1] - l64) / abs(l64):12.2e}")
    print(f"{'embeddings h':28s} {rel(h32, h64):12.2e} {rel(res['tf32x3'][0], h64):12.2e} {rel(res['simt'][0], h64):12.2e}")
    for i, k in enumerate(names):
        print(f"{'grad ' + k:28s} {rel(g32[i], g64[i]):12.2e} {rel(res['tf32x3'][2][i], g64[i]):12.2e} {rel(res['simt'][2][i], g64[i]):12.2e}")
    sys.stdout.flush()


def main():
    which = set(sys.argv[1:]) or {"c5", "c4"}
    g = torch.Generator().manual_seed(2)
    if "c5" in which:
        from test_gpu_config_sizes import _hub_graph
        n = 200_000
        ei = _hub_graph(n)
        x = torch.randn(n, 64, generator=g)
        pos = ei[:, torch.randint(0, ei.size(1), (65536,), generator=g)].contiguous()
        neg = torch.randint(0, n, pos.size(), generator=g)
        run("C5 hub slice (200k nodes, 3.5M messages, hubs of 150k / 120k edges), 64 -> 64 -> 64", x, ei, pos, neg, 64, 64, 2)
    if "c4" in which:
        data, split = synthetic_dataset("collab", seed=0)
        pos_all = split["train"]["edge"]
        pos = pos_all[torch.randint(0, pos_all.size(0), (65536,), generator=g)].t().contiguous()
        neg = torch.randint(0, data.x.size(0), pos.size(), generator=g)
        run("C4 collab shape, 128 -> 256 -> 256 -> 256", data.x, data.adj_t, pos, neg, 128, 256, 3)


if __name__ == "__main__":
    main()
