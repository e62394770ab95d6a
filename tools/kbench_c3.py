#!/usr/bin/env python
"""Micro-benchmarks at the Coauthor-Physics production shapes (BASELINE.json configs[2]): the
8415-wide first SAGEConv_updated layer is the GEMM-bound part of the path (tensor roofline), plus the C2 / collab
student distillation losses.  CUDA-event timing, L2 flushed between iterations.  Development tool.

    python tools/kbench_c3.py [gemm] [layer] [step] [kd] [student]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import linkless_link_prediction_b200 as L  # noqa: E402
from linkless_link_prediction_b200 import ops, shims  # noqa: E402
from linkless_link_prediction_b200.data import undirected_graph  # noqa: E402
from linkless_link_prediction_b200 import train_teacher_gnn as teacher  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=8, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3  # us


import json  # noqa: E402
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def report(name, us, nbytes=None, flops=None):
    s = f"{name:64s} {us:9.1f} us"
    if nbytes:
        g = nbytes / us / 1e3
        s += f"  {g:8.0f} GB/s ({100 * g / PEAK['hbm_gbs']:5.1f}% of measured HBM)"
    if flops:
        t = flops / us / 1e6
        s += f"  {t:7.1f} TFLOP/s ({100 * t / PEAK['bf16_tflops']:5.1f}% of measured bf16)"
    print(s, flush=True)


def main():
    which = set(sys.argv[1:]) or {"gemm", "layer", "step", "kd", "student"}
    n, F, H = 34493, 8415, 256
    torch.manual_seed(0)
    x32 = (torch.rand(n, F, device=dev) < 0.004).float()
    x = ops.to_compute(x32)
    if "gemm" in which:
        W = ops.to_compute(torch.randn(H, F, device=dev) * 0.01)
        W2 = ops.to_compute(torch.randn(2 * H, F, device=dev) * 0.01)
        bias = torch.randn(H, device=dev)
        bias2 = torch.randn(2 * H, device=dev)
        nb = n * F * 2 + H * F * 2 + n * H * 2
        report(f"gemm_nt x[{n},{F}] W[{H},{F}]^T + b", timeit(lambda: ops.gemm_nt(x, W, bias=bias)), nbytes=nb, flops=2 * n * F * H)
        report(f"gemm_nt x[{n},{F}] [W_l;W_r][{2 * H},{F}]^T + b (stacked)", timeit(lambda: ops.gemm_nt(x, W2, bias=bias2)),
               nbytes=n * F * 2 + 2 * H * F * 2 + 2 * n * H * 2, flops=2 * n * F * 2 * H)
        g = torch.randn(n, H, device=dev).bfloat16()
        report(f"gemm_tn g[{n},{H}]^T x[{n},{F}] (weight gradient)", timeit(lambda: ops.gemm_tn(g, x)),
               nbytes=n * (F + H) * 2 + H * F * 4, flops=2 * n * F * H)
        g2 = torch.randn(n, 2 * H, device=dev).bfloat16()
        report(f"gemm_tn [gt|g][{n},{2 * H}]^T x[{n},{F}] (stacked weight gradient)", timeit(lambda: ops.gemm_tn(g2, x)),
               nbytes=n * (F + 2 * H) * 2 + 2 * H * F * 4, flops=2 * n * F * 2 * H)
    ei = undirected_graph(n, 247962, 0, True, unique=False).to(dev)
    if "layer" in which:
        conv = L.SAGEConv_updated(F, H).to(dev)
        graph = ops.graph_of(ei, n)
        params = list(conv.parameters())
        opt = L.FusedAdam(params, lr=0.01)  # gives the parameters flat fp32 grads + bf16 working copies

        def fwd_bwd():
            y = conv(x32, ei)
            y.backward(torch.ones_like(y))

        def fwd():
            with torch.no_grad():
                conv(x32, ei)
        report("SAGEConv_updated 8415->256 forward", timeit(fwd), flops=2 * 2 * n * F * H)
        report("SAGEConv_updated 8415->256 forward + backward", timeit(fwd_bwd), flops=4 * 2 * n * F * H)
        del graph, opt
    if "step" in which:
        shims.seed_everything(0)
        data = shims.Data(x=x32, adj_t=ei)
        model = L.SAGE("coauthor-physics", F, H, H, 2, 0.5, L.SAGEConv_updated).to(dev)
        predictor = L.LinkPredictor("mlp", H, H, 1, 2, 0.5).to(dev)
        optimizer = L.FusedAdam(list(model.parameters()) + list(predictor.parameters()), lr=0.005)
        model.train(); predictor.train()
        B = 65536
        step = teacher.CapturedTrainStep(model, predictor, data, optimizer, eager_steps=2)
        pos = ei.t().contiguous()

        def one():
            perm = torch.randint(0, pos.size(0), (B,), device=dev)
            edge = pos[perm].t()
            neg = torch.randint(0, n, edge.size(), dtype=torch.long, device=dev)
            return step(edge, neg)
        us = timeit(one, iters=8, warm=4)
        report(f"C3 teacher train step (SAGEConv_updated x2 + predictor, B={B})", us, flops=6 * 2 * n * F * H)
        print(f"   => {B / us:.2f} M positive edges/s", flush=True)
    if "student" in which:
        # collab student (scripts/LLP_transductive.sh: hidden 1024, 3 layers, K = 36 context nodes, 13,110 anchors per
        # mini-batch): the predictor's first layer over [B_n*K, 1024] rows is a square, compute-bound GEMM
        M, Hs = 13110 * 36, 1024
        A = torch.randn(M, Hs, device=dev).bfloat16()
        W = (torch.randn(Hs, Hs, device=dev) * 0.03).bfloat16()
        bias = torch.randn(Hs, device=dev)
        nb = M * Hs * 2 * 2 + Hs * Hs * 2
        report(f"predictor lin1 fwd [{M},{Hs}] x [{Hs},{Hs}]^T + b, relu, dropout", timeit(lambda: ops.gemm_nt(A, W, bias=bias, relu=True, dropout_p=0.5, seed=1)),
               nbytes=nb, flops=2 * M * Hs * Hs)
        report(f"predictor lin1 dgrad [{M},{Hs}] x [{Hs},{Hs}]", timeit(lambda: ops.gemm_nt(A, W)), nbytes=nb, flops=2 * M * Hs * Hs)
        G = torch.randn(M, Hs, device=dev).bfloat16()
        report(f"predictor lin1 wgrad g[{M},{Hs}]^T z[{M},{Hs}]", timeit(lambda: ops.gemm_tn(G, A)), nbytes=nb, flops=2 * M * Hs * Hs)
        Mm = 235868
        X = torch.randn(Mm, 128, device=dev).bfloat16()
        W1 = (torch.randn(Hs, 128, device=dev) * 0.05).bfloat16()
        report(f"student MLP layer 1 [{Mm},128] -> 1024", timeit(lambda: ops.gemm_nt(X, W1, bias=bias, relu=True, dropout_p=0.5, seed=1)),
               nbytes=Mm * (128 + Hs) * 2, flops=2 * Mm * 128 * Hs)
        Hh = torch.randn(Mm, Hs, device=dev).bfloat16()
        report(f"student MLP layer 2 [{Mm},1024] -> 1024", timeit(lambda: ops.gemm_nt(Hh, W, bias=bias, relu=True, dropout_p=0.5, seed=1)),
               nbytes=Mm * 2 * Hs * 2, flops=2 * Mm * Hs * Hs)
    if "kd" in which:
        for rows, K, tag in ((2708, 12, "C2 Cora student"), (13110, 36, "collab student minibatch")):
            s = torch.rand(rows, K, device=dev); t = torch.rand(rows, K, device=dev)
            report(f"LLP_D kl_loss [{rows},{K}] fwd+grad {tag}", timeit(lambda: ops.kl_loss(s, t, 1.0)), nbytes=rows * K * 12)
            report(f"LLP_R rank_loss [{rows},{K}] ({K * (K - 1) // 2} pairs/row) fwd+grad {tag}", timeit(lambda: ops.rank_loss(s, t, 0.1)),
                   nbytes=rows * K * 12)


if __name__ == "__main__":
    main()
