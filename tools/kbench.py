#!/usr/bin/env python
"""Per-kernel micro-benchmarks at the collab (C4) sizes: CUDA-event timing on the launching stream, L2 flushed
between iterations (a 512 MB memset), achieved GB/s / TFLOP/s against MEASURED_PEAKS.json.  Development tool."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linkless_link_prediction_b200 import _native as N  # noqa: E402
from linkless_link_prediction_b200 import ops  # noqa: E402
from linkless_link_prediction_b200.data import undirected_graph  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, iters=10, warm=3):
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


def report(name, us, nbytes=None, flops=None):
    s = f"{name:58s} {us:9.1f} us"
    if nbytes:
        g = nbytes / us / 1e3
        s += f"  {g:8.0f} GB/s ({100 * g / PEAK['hbm_gbs']:5.1f}% of measured HBM)"
    if flops:
        t = flops / us / 1e6
        s += f"  {t:7.1f} TFLOP/s ({100 * t / PEAK['bf16_tflops']:5.1f}% of measured bf16)"
    print(s, flush=True)


def main():
    which = set(sys.argv[1:]) or {"spmm", "gemm", "misc"}
    n, E = 235868, 2358104
    ei = undirected_graph(n, 1179052, 0, True, unique=False).to(dev)
    g = ops.Graph(ei, n)
    if "spmmsweep" in which:
        for variant in (0, 1, 2, 3):  # 0 = default (groups of 4 for 512-byte rows), 3 = groups of 8
            N.load().llp_set_tuning(0, variant)
            for dt, F in ((torch.bfloat16, 256), (torch.bfloat16, 128), (torch.float32, 256)):
                x = torch.randn(n, F, device=dev).to(dt)
                s = x.element_size()
                nb = E * F * s + n * F * s + 4 * E + 4 * (n + 1)
                report(f"variant {variant}: spmm fwd {dt} F={F}", timeit(lambda: g.spmm(x)), nbytes=nb)
                report(f"variant {variant}: spmm bwd {dt} F={F}", timeit(lambda: g.spmm(x, transpose=True)), nbytes=nb)
        N.load().llp_set_tuning(0, 0)
    if "spmmab" in which:   # row-run kernel (llp_set_tuning(3, 1)) vs streaming kernel (3, 0), per occupancy / group variant
        lib = N.load()
        for dt, F in ((torch.bfloat16, 256), (torch.bfloat16, 128), (torch.float32, 256), (torch.float32, 128)):
            x = torch.randn(n, F, device=dev).to(dt)
            s = x.element_size()
            nb = E * F * s + n * F * s + 4 * E + 4 * (n + 1)
            lib.llp_set_tuning(3, 1); lib.llp_set_tuning(0, 0)
            ref_f, ref_b = g.spmm(x), g.spmm(x, transpose=True)
            for stream, pf in ((1, 0), (2, 0)):
                for variant in (0, 4, 5, 6):
                    lib.llp_set_tuning(3, stream); lib.llp_set_tuning(0, variant); lib.llp_set_tuning(4, 1 if pf else 0)
                    same = bool(torch.equal(g.spmm(x), ref_f) and torch.equal(g.spmm(x, transpose=True), ref_b))
                    tag = f"{ {1: 'row-run', 2: 'stream '}[stream]}{' +L2 prefetch' if pf else ''} variant {variant} {str(dt)[6:]} F={F} same={same}"
                    report(f"{tag} fwd", timeit(lambda: g.spmm(x)), nbytes=nb)
                    report(f"{tag} bwd", timeit(lambda: g.spmm(x, transpose=True)), nbytes=nb + 4 * n)
        lib.llp_set_tuning(4, 0)
        lib.llp_set_tuning(3, 0); lib.llp_set_tuning(0, 0)
    if "spmmexp" in which:
        N.load().llp_set_tuning(0, 0)
        x = torch.randn(n, 256, device=dev).bfloat16()
        nb = E * 512 + n * 512 + 4 * E + 4 * (n + 1)
        for div in (1, 2, 4, 8):
            N.load().llp_set_tuning(1, div)
            report(f"chunk_div={div} (1/{div} of the chunks) bf16 F=256", timeit(lambda: g.spmm(x)), nbytes=nb / div)
        N.load().llp_set_tuning(1, 1)
        N.load().llp_set_tuning(2, 1)
        report("sequential sources (e mod N) bf16 F=256", timeit(lambda: g.spmm(x)), nbytes=nb)
        N.load().llp_set_tuning(2, 0)
        N.load().llp_set_tuning(0, 0)
    if "spmm" in which:
        for dt in (torch.bfloat16, torch.float32):
            for F in (128, 256):
                x = torch.randn(n, F, device=dev).to(dt)
                s = x.element_size()
                nb = E * F * s + n * F * s + 4 * E + 4 * (n + 1)
                report(f"spmm fwd  {dt} F={F}", timeit(lambda: g.spmm(x)), nbytes=nb)
                report(f"spmm bwd  {dt} F={F}", timeit(lambda: g.spmm(x, transpose=True)), nbytes=nb + 4 * n)
    if "gemm" in which:
        H = 256
        for M, K1, K2, kw, tag in ((n, 128, 128, dict(relu=True, dropout_p=0.5, seed=1), "L1 fwd relu+drop"),
                                   (n, 256, 256, dict(relu=True, dropout_p=0.5, seed=1), "L2 fwd relu+drop"),
                                   (n, 256, 256, dict(), "L3 fwd / dgrad"),
                                   (n, 256, 256, dict(relu=True), "L2 fwd relu only"),
                                   (131072, 256, 0, dict(relu=True, dropout_p=0.5, seed=1), "pred fwd relu+drop"),
                                   (131072, 256, 0, dict(), "pred dgrad")):
            A1 = torch.randn(M, K1, device=dev).bfloat16(); B1 = torch.randn(H, K1, device=dev).bfloat16()
            A2 = torch.randn(M, K2, device=dev).bfloat16() if K2 else None
            B2 = torch.randn(H, K2, device=dev).bfloat16() if K2 else None
            bias = torch.randn(H, device=dev)
            us = timeit(lambda: ops.gemm_nt(A1, B1, A2, B2, bias=bias, **kw))
            report(f"gemm_nt M={M} K={K1}+{K2} {tag}", us, nbytes=M * (K1 + K2) * 2 + M * H * 2, flops=2 * M * H * (K1 + K2))
        for M, N1, N2 in ((n, 256, 256), (n, 256, 128), (131072, 256, 256)):
            A = torch.randn(M, N1, device=dev).bfloat16(); B = torch.randn(M, N2, device=dev).bfloat16()
            report(f"gemm_tn M={M} {N1}x{N2}", timeit(lambda: ops.gemm_tn(A, B)), nbytes=M * (N1 + N2) * 2, flops=2 * M * N1 * N2)
    if "wgrad" in which or "gemm" in which:
        for M, N1, n2a, n2b, tag in ((n, 256, 256, 256, "L2/L3"), (n, 256, 128, 128, "L1"), (131072, 256, 256, 0, "predictor")):
            G = torch.randn(M, N1, device=dev).bfloat16(); A = torch.randn(M, n2a, device=dev).bfloat16()
            B = torch.randn(M, n2b, device=dev).bfloat16() if n2b else None
            Wa = torch.nn.Parameter(torch.zeros(N1, n2a, device=dev)); Wa.grad = torch.zeros_like(Wa)
            Wb = torch.nn.Parameter(torch.zeros(N1, max(n2b, 1), device=dev)); Wb.grad = torch.zeros_like(Wb)
            bp = torch.nn.Parameter(torch.zeros(N1, device=dev)); bp.grad = torch.zeros_like(bp)
            us = timeit(lambda: ops.wgrad(G, A, Wa, B, Wb if n2b else None, bias=bp))
            report(f"wgrad fused M={M} {N1}x({n2a}+{n2b})+bias {tag}", us, nbytes=M * (N1 + n2a + n2b) * 2,
                   flops=2 * M * N1 * (n2a + n2b))
    if "tf32" in which:
        # fp32-parity mode: 3xTF32 tensor-core GEMMs (gemm_tf32.cu) vs the CUDA-core fp32 kernels, C4 layer shapes
        H = 256
        tf_peak = PEAK["bf16_tflops"] / 2.0   # dense TF32 runs at half the bf16 tensor rate; 3 MMAs per product => / 3 of that
        for M, K1, K2, kw, tag in ((n, 128, 128, dict(relu=True, dropout_p=0.5, seed=1), "L1 fwd relu+drop"),
                                   (n, 256, 256, dict(relu=True, dropout_p=0.5, seed=1), "L2 fwd relu+drop"),
                                   (n, 256, 256, dict(), "L3 fwd / dgrad"),
                                   (131072, 256, 0, dict(relu=True, dropout_p=0.5, seed=1), "pred fwd relu+drop"),
                                   (34493, 8415, 0, dict(), "C3 layer 1 (N=256)")):
            A1 = torch.randn(M, K1, device=dev); B1 = torch.randn(H, K1, device=dev)
            A2 = torch.randn(M, K2, device=dev) if K2 else None
            B2 = torch.randn(H, K2, device=dev) if K2 else None
            bias = torch.randn(H, device=dev)
            A1, B1 = ops.cast2d(A1, torch.float32), ops.cast2d(B1, torch.float32)
            fl = 2 * M * H * (K1 + K2)
            nb = M * (K1 + K2) * 4 + M * H * 4
            for bname, be in (("tf32x3", N.GEMM_TF32X3), ("simt", N.GEMM_SIMT)):
                us = timeit(lambda: ops.gemm_nt(A1, B1, A2, B2, bias=bias, backend=be, **kw), iters=5)
                t = fl / us / 1e6
                print(f"gemm_nt fp32 [{bname}] M={M} K={K1}+{K2} {tag:22s} {us:9.1f} us  {t:7.1f} fp32-TFLOP/s "
                      f"({100 * 3 * t / tf_peak:5.1f}% of the TF32 tensor peak counting 3 MMAs)  {nb / us / 1e3:7.0f} GB/s", flush=True)
        for M, N1, N2 in ((n, 256, 256), (n, 256, 128), (131072, 256, 256)):
            A = torch.randn(M, N1, device=dev); B = torch.randn(M, N2, device=dev)
            fl = 2 * M * N1 * N2
            for bname, be in (("tf32x3", N.GEMM_TF32X3), ("simt", N.GEMM_SIMT)):
                us = timeit(lambda: ops.gemm_tn(A, B, backend=be), iters=5)
                t = fl / us / 1e6
                print(f"gemm_tn fp32 [{bname}] M={M} {N1}x{N2} {us:9.1f} us  {t:7.1f} fp32-TFLOP/s "
                      f"({100 * 3 * t / tf_peak:5.1f}% of the TF32 tensor peak counting 3 MMAs)", flush=True)
    if "tall" in which:   # tall-tile NT kernel (opt-in, llp_set_tuning(20, 3)) against the streaming kernel (the default)
        lib = N.load()
        H = 256
        for M, K1, K2, kw, tag in ((n, 256, 256, dict(relu=True, dropout_p=0.5, seed=1), "L2 fwd relu+drop"),
                                   (n, 256, 256, dict(), "L3 fwd / dgrad"), (n + 77, 256, 256, dict(relu=True), "odd M"),
                                   (n, 320, 0, dict(), "single operand K=320"), (n, 256, 256, dict(gate=True), "dgrad with gate")):
            A1 = torch.randn(M, K1, device=dev).bfloat16(); B1 = torch.randn(H, K1, device=dev).bfloat16()
            A2 = torch.randn(M, K2, device=dev).bfloat16() if K2 else None
            B2 = torch.randn(H, K2, device=dev).bfloat16() if K2 else None
            bias = torch.randn(H, device=dev)
            kw = dict(kw)
            if kw.pop("gate", False):
                kw.update(gate=torch.randn(M, H, device=dev).bfloat16(), gate_scale=2.0)
            outs = {}
            for vtag, knob in (("tall", 3), ("streaming", 0)):
                lib.llp_set_tuning(20, knob)
                outs[vtag] = ops.gemm_nt(A1, B1, A2, B2, bias=bias, **kw)
                report(f"gemm_nt [{vtag}] M={M} K={K1}+{K2} {tag}",
                       timeit(lambda: ops.gemm_nt(A1, B1, A2, B2, bias=bias, **kw)), nbytes=M * (K1 + K2) * 2 + M * H * 2,
                       flops=2 * M * H * (K1 + K2))
            lib.llp_set_tuning(20, 0)
            print(f"   bit-identical: {bool(torch.equal(outs['tall'], outs['streaming']))}", flush=True)
    if "gemmexp" in which:
        import ctypes
        lib = N.load()
        H = 256
        for M, K1, K2, kw, tag in ((n, 256, 256, dict(relu=True, dropout_p=0.5, seed=1), "L2 fwd relu+drop"),
                                   (n, 256, 256, dict(), "L3 fwd"), (131072, 256, 0, dict(relu=True, dropout_p=0.5, seed=1), "pred fwd"),
                                   (n, 128, 128, dict(relu=True, dropout_p=0.5, seed=1), "L1 fwd relu+drop"),
                                   (n, 128, 128, dict(), "L1-shaped, plain epilogue"), (131072, 256, 0, dict(), "pred dgrad")):
            A1 = torch.randn(M, K1, device=dev).bfloat16(); B1 = torch.randn(H, K1, device=dev).bfloat16()
            A2 = torch.randn(M, K2, device=dev).bfloat16() if K2 else None
            B2 = torch.randn(H, K2, device=dev).bfloat16() if K2 else None
            bias = torch.randn(H, device=dev)
            for vtag, knobs in (("auto", {}), ("CTA pair (opt-in)", {20: 2}), ("CTA pair, no MMAs (load pipeline only)", {20: 2, 11: 1}),
                                ("streaming", {16: 1})):
                for k_ in (11, 16, 18, 20):
                    lib.llp_set_tuning(k_, knobs.get(k_, 0))
                report(f"gemm_nt [{vtag}] M={M} K={K1}+{K2} {tag}",
                       timeit(lambda: ops.gemm_nt(A1, B1, A2, B2, bias=bias, **kw)), nbytes=M * (K1 + K2) * 2 + M * H * 2)
                lib.llp_set_tuning(15, 1)
                flush.zero_()
                ops.gemm_nt(A1, B1, A2, B2, bias=bias, **kw)
                buf = (ctypes.c_int64 * (148 * 4))()
                lib.llp_debug_read(buf, 148 * 4)
                lib.llp_set_tuning(15, 0)
                act = [i for i in range(148) if buf[4 * i] > 0]   # (the pair kernel reports from leader CTAs only)
                tot = sorted(buf[4 * i] for i in act); wf = sorted(buf[4 * i + 1] for i in act); wa = sorted(buf[4 * i + 2] for i in act)
                h_ = len(act) // 2
                print(f"   MMA issue loop (ns) min/med/max {tot[0]}/{tot[h_]}/{tot[-1]}; waiting for operands med/max {wf[h_]}/{wf[-1]}; "
                      f"waiting for a free accumulator med/max {wa[h_]}/{wa[-1]} ({len(act)} issuing CTAs)", flush=True)
                for i in range(148 * 4):
                    buf[i] = 0
            for k_ in (11, 16, 18, 20):
                lib.llp_set_tuning(k_, 0)
    if "edgemlp" in which:
        import ctypes
        import linkless_link_prediction_b200 as L
        lib = N.load()
        H, M = 256, 131072
        hN = torch.randn(n, H, device=dev).bfloat16()
        u = torch.randint(0, n, (M,), device=dev); v = torch.randint(0, n, (M,), device=dev)
        pred = L.LinkPredictor("mlp", H, H, 1, 2, 0.5).to(dev)
        for tag, train in (("eval (scores only)", False), ("train forward (z, y, prob; dropout)", True)):
            pred.train(train)
            hh = hN.clone().requires_grad_(train)
            def fwd():
                with torch.set_grad_enabled(train):
                    return pred.score(hh, u, v)
            nb = M * 2 * H * 2 + (M * 2 * H * 2 if train else 0) + 4 * M
            report(f"fused edge scorer M={M} H={H} {tag}", timeit(fwd), nbytes=nb, flops=2 * M * H * H)
            lib.llp_set_tuning(15, 1)
            flush.zero_()
            fwd()
            buf = (ctypes.c_int64 * (148 * 4))()
            lib.llp_debug_read(buf, 148 * 4)
            lib.llp_set_tuning(15, 0)
            tot = sorted(buf[4 * i] for i in range(148)); wf = sorted(buf[4 * i + 1] for i in range(148)); wa = sorted(buf[4 * i + 2] for i in range(148))
            print(f"   MMA issue loop (ns) min/med/max {tot[0]}/{tot[74]}/{tot[-1]}; waiting for operands med/max {wf[74]}/{wf[-1]}; "
                  f"waiting for a free accumulator med/max {wa[74]}/{wa[-1]}", flush=True)
    if "wgradexp" in which:
        lib = N.load()
        M, N1 = n, 256
        G = torch.randn(M, N1, device=dev).bfloat16(); A = torch.randn(M, 256, device=dev).bfloat16()
        B = torch.randn(M, 256, device=dev).bfloat16()
        Wa = torch.nn.Parameter(torch.zeros(N1, 256, device=dev)); Wa.grad = torch.zeros_like(Wa)
        Wb = torch.nn.Parameter(torch.zeros(N1, 256, device=dev)); Wb.grad = torch.zeros_like(Wb)
        bp = torch.nn.Parameter(torch.zeros(N1, device=dev)); bp.grad = torch.zeros_like(bp)
        nb = M * 768 * 2
        for tag, knobs, bias in (("default", {}, True), ("no bias warp", {}, False), ("stages=3", {10: 3}, True),
                                 ("stages=2", {10: 2}, True), ("skip mma", {11: 1}, True), ("skip mma, no bias", {11: 1}, False),
                                 ("BK=64 (2 stages)", {12: 64}, True), ("BK=64 skip mma no bias", {12: 64, 11: 1}, False),
                                 ("wide un-swizzled boxes, skip mma, no bias", {13: 1}, False),
                                 ("wide boxes BK=64", {13: 1, 12: 64}, False),
                                 ("strided k-blocks", {14: 1}, True), ("strided k-blocks BK=64", {14: 1, 12: 64}, True),
                                 ("strided k-blocks, 3 stages", {14: 1, 10: 3}, True)):
            for k in (10, 11, 12, 13, 14):
                lib.llp_set_tuning(k, knobs.get(k, 0))
            report(f"wgrad 256x(256+256) {tag}", timeit(lambda: ops.wgrad(G, A, Wa, B, Wb, bias=bp if bias else None)), nbytes=nb)
        for k in (10, 11, 12, 13, 14):
            lib.llp_set_tuning(k, 0)
        # phase stamps (globaltimer, ns) of every CTA of one launch: main loop vs epilogue
        import ctypes
        lib.llp_set_tuning(15, 1)
        ops.wgrad(G, A, Wa, B, Wb, bias=bp)
        buf = (ctypes.c_int64 * (148 * 4))()
        lib.llp_debug_read(buf, 148 * 4)
        lib.llp_set_tuning(15, 0)
        st = [(buf[4 * i], buf[4 * i + 1], buf[4 * i + 2]) for i in range(148)]
        t0 = min(s_[0] for s_ in st)
        main = sorted(s_[1] - s_[0] for s_ in st); epi = sorted(s_[2] - s_[1] for s_ in st)
        print(f"wgrad phases (ns): start skew {max(s_[0] for s_ in st) - t0}, main loop min/med/max {main[0]}/{main[74]}/{main[-1]}, "
              f"epilogue min/med/max {epi[0]}/{epi[74]}/{epi[-1]}, last CTA done at {max(s_[2] for s_ in st) - t0}", flush=True)
        # no duplicated reads: a single N1 tile (G is [M,128]) => every byte is fetched by exactly one CTA
        G1 = torch.randn(M, 128, device=dev).bfloat16()
        Wa1 = torch.nn.Parameter(torch.zeros(128, 256, device=dev)); Wa1.grad = torch.zeros_like(Wa1)
        Wb1 = torch.nn.Parameter(torch.zeros(128, 256, device=dev)); Wb1.grad = torch.zeros_like(Wb1)
        for tag, knobs in (("default", {}), ("skip mma", {11: 1}), ("strided k-blocks", {14: 1})):
            for k in (10, 11, 12, 13, 14):
                lib.llp_set_tuning(k, knobs.get(k, 0))
            report(f"wgrad 128x(256+256) single N1 tile, {tag}", timeit(lambda: ops.wgrad(G1, A, Wa1, B, Wb1)), nbytes=M * 640 * 2)
        for k in (10, 11, 12, 13, 14):
            lib.llp_set_tuning(k, 0)
    if "misc" in which:
        H = 256
        a = torch.randn(n, H, device=dev).bfloat16(); b = torch.randn(n, H, device=dev).bfloat16()
        report("gate [N,256] bf16", timeit(lambda: ops.gate(a, b, 2.0)), nbytes=3 * n * H * 2)
        report("colsum [N,256] bf16", timeit(lambda: ops.colsum(a)), nbytes=n * H * 2)
        xf = torch.randn(n, 128, device=dev)
        report("cast2d fp32->bf16 [N,128]", timeit(lambda: ops.cast2d(xf, torch.bfloat16)), nbytes=n * 128 * 6)
        M = 131072
        u = torch.randint(0, n, (M,), device=dev); v = torch.randint(0, n, (M,), device=dev)
        report("edge_hadamard M=131072", timeit(lambda: ops.HadamardFn.apply(a, u, v)), nbytes=3 * M * H * 2)
        z = torch.randn(M, H, device=dev).bfloat16()
        hh = a.clone().requires_grad_(True)
        def hb():
            zz = ops.HadamardFn.apply(hh, u, v)
            zz.backward(z)
            hh.grad = None
        report("edge_hadamard fwd + plan (sort) + gather-reduce bwd", timeit(hb), nbytes=(3 + 3) * M * H * 2 + n * H * 10)
        plan = ops.EdgePlan(u, v, n)
        def hb2():
            zz = ops.HadamardFn.apply(hh, u, v, plan)
            zz.backward(z)
            hh.grad = None
        report("edge_hadamard fwd + gather-reduce bwd (plan prebuilt)", timeit(hb2), nbytes=(3 + 3) * M * H * 2 + n * H * 2)
        report("edge plan (incidence sort) M=131072", timeit(lambda: ops.EdgePlan(u, v, n)))
        w = torch.randn(1, H, device=dev); bb = torch.randn(1, device=dev)
        report("score_head fwd M=131072", timeit(lambda: ops.ScoreHeadFn.apply(z, w, bb)), nbytes=M * H * 2)


if __name__ == "__main__":
    main()
