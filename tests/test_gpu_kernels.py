"""Parity tests proper: every CUDA kernel, called through the C-ABI (ctypes), against the CPU oracle on the same
seeded inputs.  Tolerances are the ones BASELINE.json's north_star states: bit-exact for CSR construction, sampled
indices and Hits@K counts; 1e-5 relative in fp32 mode; 2e-2 in bf16 mode."""
import itertools
import math

import numpy as np
import pytest
import torch

from linkless_link_prediction_b200 import _native as N
from linkless_link_prediction_b200 import ops
from oracle import llp_oracle as O

pytestmark = pytest.mark.gpu

FP32 = dict(rtol=1e-5, atol=1e-6)
BF16 = dict(rtol=2e-2, atol=2e-2)


def rand_graph(n, e, seed, hub=None):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, max(n - 3, 1), (e,), generator=g)  # the last nodes receive nothing (empty rows)
    if hub is not None:
        dst[: hub] = 1  # one destination with a huge in-degree
    return torch.stack([src, dst])


# ------------------------------------------------------------------------------------------------
# graph structure
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,e,hub", [(50, 0, None), (7, 20, None), (1000, 5000, None), (300, 9000, 4000), (5, 1, None)])
def test_csr_build_bit_exact(cuda, n, e, hub):
    ei = rand_graph(n, e, 1, hub)
    g = ops.Graph(ei.to(cuda), n)
    for by, (rp, col, perm) in (("dst", (g.rowptr, g.col, g.perm)), ("src", (g.t_rowptr, g.t_col, g.t_perm))):
        o_rp, o_col, o_perm = O.csr_build(ei, n, by)
        assert torch.equal(rp.cpu(), o_rp)
        assert torch.equal(col.cpu()[:e], o_col)
        assert torch.equal(perm.cpu()[:e], o_perm)
    deg = torch.bincount(ei[1], minlength=n).clamp(min=1).float()
    assert torch.equal(g.inv_deg.cpu(), 1.0 / deg)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,e,f,hub", [(64, 0, 16, None), (200, 1500, 128, None), (500, 4000, 256, None),
                                       (300, 12000, 256, 5000), (120, 700, 1433, None), (90, 500, 7, None),
                                       (2000, 30000, 64, 1500), (3000, 50000, 512, 3000), (50, 3000, 8, 2000),
                                       (1, 1, 128, None), (700, 128 * 5, 32, None)])
def test_spmm_forward_and_transpose(cuda, dtype, n, e, f, hub):
    ei = rand_graph(n, e, 2, hub)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(3))
    xq = x.to(dtype).float()  # what the kernel actually reads
    g = ops.Graph(ei.to(cuda), n)
    xc = ops.cast2d(x.to(cuda), dtype)
    tol = FP32 if dtype == torch.float32 else BF16
    out = g.spmm(xc).float().cpu()
    torch.testing.assert_close(out, O.mean_aggregate(xq, ei, n), **tol)
    # transpose: gx[s] = sum_{s->d} g[d] / deg(d)  == autograd of the oracle's mean aggregation
    xo = xq.clone().requires_grad_(True)
    gy = torch.randn(n, f, generator=torch.Generator().manual_seed(4)).to(dtype).float()
    O.mean_aggregate(xo, ei, n).backward(gy)
    gx = g.spmm(ops.cast2d(gy.to(cuda), dtype), transpose=True).float().cpu()
    torch.testing.assert_close(gx, xo.grad, **(FP32 if dtype == torch.float32 else dict(rtol=2e-2, atol=6e-2)))


def test_spmm_fp32_is_bit_exact_in_edge_order(cuda):
    # no hub rows: the kernel sums each row in CSR (= original edge) order like index_add_ on the CPU
    ei = rand_graph(400, 3000, 5)
    x = torch.randn(400, 128, generator=torch.Generator().manual_seed(6))
    g = ops.Graph(ei.to(cuda), 400)
    out = g.spmm(x.to(cuda)).cpu()
    rp, col, _ = O.csr_build(ei, 400)
    assert torch.equal(out, O.spmm_csr(rp, col, x, True))


def test_spmm_collab_size_properties(cuda):
    # BASELINE config C4 size: checked through size-independent properties (no CPU oracle at this size)
    n, f = 235868, 256
    from linkless_link_prediction_b200.data import undirected_graph
    ei = undirected_graph(n, 1179052, 0, True, unique=False).to(cuda)
    g = ops.Graph(ei, n)
    deg = (g.rowptr[1:] - g.rowptr[:-1]).float()
    ones = torch.ones(n, f, dtype=torch.bfloat16, device=cuda)
    out = g.spmm(ones).float()
    assert torch.equal(out[deg > 0], torch.ones_like(out[deg > 0]))       # mean of ones
    assert torch.equal(out[deg == 0], torch.zeros_like(out[deg == 0]))    # isolated rows are zero-filled
    x = torch.randn(n, f, device=cuda)
    a, b = g.spmm(x), g.spmm(2 * x)
    torch.testing.assert_close(b, 2 * a, rtol=1e-6, atol=1e-6)            # linearity
    # <A x, y> == <x, A^T y>  (forward vs transpose kernels, fp32)
    y = torch.randn(n, f, device=cuda)
    lhs = (a.double() * y.double()).sum()
    rhs = (x.double() * g.spmm(y, transpose=True).double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


# ------------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------------
def _ref_gemm(A1, B1, A2, B2, bias, addend, relu, gate, gate_scale):
    D = A1.double() @ B1.double().t()
    if A2 is not None:
        D = D + A2.double() @ B2.double().t()
    if bias is not None:
        D = D + bias.double()
    if addend is not None:
        D = D + addend.double()
    if relu:
        D = D.clamp(min=0)
    if gate is not None:
        D = torch.where(gate.double() > 0, D * gate_scale, torch.zeros_like(D))
    return D.float()


@pytest.mark.parametrize("backend", [N.GEMM_SIMT, N.GEMM_TCGEN05, N.GEMM_TF32X3], ids=["simt", "tcgen05", "tf32x3"])
@pytest.mark.parametrize("M,Nn,K1,K2", [(1, 8, 8, 0), (127, 64, 64, 0), (128, 256, 256, 256), (300, 200, 100, 36),
                                        (1000, 256, 1433, 0), (5000, 256, 128, 128), (129, 16, 520, 0), (2048, 1, 256, 0)])
def test_gemm_nt(cuda, backend, M, Nn, K1, K2):
    dtype = torch.bfloat16 if backend == N.GEMM_TCGEN05 else torch.float32
    g = torch.Generator().manual_seed(M + Nn + K1)
    r = lambda *s: torch.randn(*s, generator=g)
    A1, B1 = r(M, K1).to(dtype), r(Nn, K1).to(dtype)
    A2, B2 = (r(M, K2).to(dtype), r(Nn, K2).to(dtype)) if K2 else (None, None)
    bias, addend, gate = r(Nn), r(M, Nn), r(M, Nn)  # addend / gate are read in the OUTPUT dtype (fp32 here)
    c = lambda t: None if t is None else ops.cast2d(t.to(cuda), t.dtype)  # row-padded device copy
    for kw in (dict(), dict(bias=bias, relu=True), dict(bias=bias, addend=addend), dict(gate=gate, gate_scale=2.0)):
        dev_kw = {k: (v.to(cuda) if k == "bias" else c(v) if torch.is_tensor(v) else v) for k, v in kw.items()}
        D = ops.gemm_nt(c(A1), c(B1), c(A2), c(B2), backend=backend, out_dtype=torch.float32, **dev_kw).cpu()
        ref = _ref_gemm(A1, B1, A2, B2, kw.get("bias"), kw.get("addend"), kw.get("relu", False), kw.get("gate"),
                        kw.get("gate_scale", 1.0))
        scale = math.sqrt(K1 + K2)
        tol = dict(rtol=1e-5, atol=2e-5 * scale) if dtype == torch.float32 else dict(rtol=1e-3, atol=1e-3 * scale)
        torch.testing.assert_close(D, ref, **tol)
    # storage-dtype output path
    D = ops.gemm_nt(c(A1), c(B1), bias=bias.to(cuda), relu=True, backend=backend).float().cpu()
    ref = _ref_gemm(A1, B1, None, None, bias, None, True, None, 1.0)
    torch.testing.assert_close(D, ref, **(dict(rtol=1e-5, atol=2e-5 * math.sqrt(K1)) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2 * math.sqrt(K1))))


@pytest.mark.parametrize("M,Nn,K1,K2", [(40000, 256, 256, 256), (38011, 256, 256, 0), (50001, 200, 128, 128),
                                        (45000, 128, 256, 256), (39999, 256, 192, 64), (41003, 200, 320, 192),
                                        (75777, 256, 512, 0)])
def test_gemm_nt_resident_weights(cuda, M, Nn, K1, K2):
    """Tall problems take the resident-weight tcgen05 kernels (K <= 256: weights loaded into shared memory once per CTA;
    K <= 512: once per CTA PAIR, tcgen05.mma.cta_group::2): same results as the streaming kernel (bit-identical: same MMA
    order per output element) and as the fp64 reference."""
    g = torch.Generator().manual_seed(M + Nn + K1)
    r = lambda *s: torch.randn(*s, generator=g)
    A1, B1 = r(M, K1).bfloat16(), r(Nn, K1).bfloat16()
    A2, B2 = (r(M, K2).bfloat16(), r(Nn, K2).bfloat16()) if K2 else (None, None)
    bias, gate = r(Nn), r(M, Nn).bfloat16()
    c = lambda t: None if t is None else ops.cast2d(t.to(cuda), t.dtype)
    dA1, dB1, dA2, dB2, dgate = c(A1), c(B1), c(A2), c(B2), c(gate)
    lib = N.load()
    outs = []
    for streaming in (0, 1):
        lib.llp_set_tuning(16, streaming)   # 1 = force the streaming kernel
        lib.llp_set_tuning(20, 0 if streaming else 2)   # 2 = allow the (opt-in) CTA-pair kernel for K <= 512
        try:
            plain = ops.gemm_nt(dA1, dB1, dA2, dB2, bias=bias.to(cuda), relu=True, out_dtype=torch.float32,
                                backend=N.GEMM_TCGEN05)
            gated = ops.gemm_nt(dA1, dB1, dA2, dB2, gate=dgate, gate_scale=2.0, backend=N.GEMM_TCGEN05)
            drop = ops.gemm_nt(dA1, dB1, dA2, dB2, relu=True, dropout_p=0.5, seed=5, offset=3, backend=N.GEMM_TCGEN05)
        finally:
            lib.llp_set_tuning(16, 0)
            lib.llp_set_tuning(20, 0)
        outs.append((plain, gated, drop))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    ref = _ref_gemm(A1, B1, A2, B2, bias, None, True, None, 1.0)
    torch.testing.assert_close(outs[0][0].cpu(), ref, rtol=1e-3, atol=1e-3 * math.sqrt(K1 + K2))


@pytest.mark.parametrize("backend", [N.GEMM_SIMT, N.GEMM_TCGEN05, N.GEMM_TF32X3], ids=["simt", "tcgen05", "tf32x3"])
@pytest.mark.parametrize("M,N1,N2", [(64, 8, 8), (1000, 256, 256), (5000, 256, 1433), (333, 200, 36), (20000, 256, 512),
                                     (100, 1, 256)])
def test_gemm_tn(cuda, backend, M, N1, N2):
    dtype = torch.bfloat16 if backend == N.GEMM_TCGEN05 else torch.float32
    if backend == N.GEMM_TCGEN05 and N1 % 8:
        pytest.skip("TMA needs 16-byte rows")
    g = torch.Generator().manual_seed(M + N1)
    A, B = torch.randn(M, N1, generator=g).to(dtype), torch.randn(M, N2, generator=g).to(dtype)
    D = ops.gemm_tn(ops.cast2d(A.to(cuda), dtype), ops.cast2d(B.to(cuda), dtype), backend=backend).cpu()
    ref = (A.double().t() @ B.double()).float()
    tol = dict(rtol=1e-5, atol=2e-5 * math.sqrt(M)) if dtype == torch.float32 else dict(rtol=1e-3, atol=1e-3 * math.sqrt(M))
    torch.testing.assert_close(D, ref, **tol)


@pytest.mark.parametrize("backend", [N.GEMM_SIMT, N.GEMM_TCGEN05, N.GEMM_TF32X3], ids=["simt_fp32", "tcgen05_bf16", "tf32x3_fp32"])
@pytest.mark.parametrize("M,N1,n2a,n2b,bias", [(1, 8, 8, 0, True), (77, 64, 128, 128, True), (1000, 256, 256, 256, True),
                                              (5000, 256, 128, 128, False), (4099, 200, 136, 72, True),
                                              (30000, 256, 256, 0, True), (3000, 384, 64, 256, True),
                                              (700, 256, 1440, 0, True)])
def test_wgrad_fused(cuda, backend, M, N1, n2a, n2b, bias):
    """llp_wgrad: dWa = G^T A, dWb = G^T B, dbias = colsum(G) in one pass, fresh outputs and accumulation into
    existing fp32 .grad buffers (autograd of lin_l(agg) + lin_r(x): two mm's and a sum(0))."""
    dtype = torch.bfloat16 if backend == N.GEMM_TCGEN05 else torch.float32
    if backend == N.GEMM_TCGEN05 and max(n2a, n2b) > 256:
        backend = N.GEMM_AUTO  # wider than the fused kernel's TMEM tile: the library takes the separate tcgen05 kernels
    gen = torch.Generator().manual_seed(M + N1 + n2a)
    G = torch.randn(M, N1, generator=gen).to(dtype)
    A = torch.randn(M, n2a, generator=gen).to(dtype)
    B = torch.randn(M, n2b, generator=gen).to(dtype) if n2b else None
    Gd, Ad = ops.cast2d(G.to(cuda), dtype), ops.cast2d(A.to(cuda), dtype)
    Bd = ops.cast2d(B.to(cuda), dtype) if n2b else None
    Wa = torch.nn.Parameter(torch.zeros(N1, n2a, device=cuda))
    Wb = torch.nn.Parameter(torch.zeros(N1, n2b, device=cuda)) if n2b else None
    bp = torch.nn.Parameter(torch.zeros(N1, device=cuda)) if bias else None
    tol = dict(rtol=1e-5, atol=2e-5 * math.sqrt(M)) if dtype == torch.float32 else dict(rtol=1e-3, atol=1e-3 * math.sqrt(M))
    ref_a = (G.double().t() @ A.double()).float()
    ref_b = (G.double().t() @ B.double()).float() if n2b else None
    ref_bias = G.double().sum(0).float()
    dWa, dWb, db = ops.wgrad(Gd, Ad, Wa, Bd, Wb, bias=bp, backend=backend)   # no .grad yet: fresh outputs
    torch.testing.assert_close(dWa.cpu(), ref_a, **tol)
    if n2b:
        torch.testing.assert_close(dWb.cpu(), ref_b, **tol)
    if bias:
        torch.testing.assert_close(db.cpu(), ref_bias, **tol)
    Wa.grad = torch.full_like(Wa, 1.0)
    if n2b:
        Wb.grad = torch.full_like(Wb, 2.0)
    if bias:
        bp.grad = torch.full_like(bp, 3.0)
    assert ops.wgrad(Gd, Ad, Wa, Bd, Wb, bias=bp, backend=backend) == (None, None, None)  # accumulated in place
    torch.testing.assert_close(Wa.grad.cpu(), ref_a + 1.0, **tol)
    if n2b:
        torch.testing.assert_close(Wb.grad.cpu(), ref_b + 2.0, **tol)
    if bias:
        torch.testing.assert_close(bp.grad.cpu(), ref_bias + 3.0, **tol)
    # deterministic: same bits on a second run
    Wa.grad = None
    again = ops.wgrad(Gd, Ad, Wa, Bd, Wb if n2b else None, bias=None, backend=backend)[0]
    assert torch.equal(again, dWa)


@pytest.mark.parametrize("M,Nn,K", [(4096, 256, 512), (1000, 200, 1433), (20000, 256, 8415)])
def test_tf32x3_is_fp32_grade(cuda, M, Nn, K):
    """The fp32-parity mode runs its GEMMs on the tensor cores as 3xTF32 split accumulation (tcgen05.mma.kind::tf32).
    Against an fp64 reference its error must be of the size of a true-fp32 FFMA GEMM's (the CUDA-core kernel, i.e. what
    the reference's cuBLAS SGEMM delivers), NOT of a single TF32 product (~1e-3 relative): bounded by 2x the fp32
    kernel's error + 1e-6 of the result scale, NT and TN."""
    g = torch.Generator().manual_seed(K)
    A, B = torch.randn(M, K, generator=g), torch.randn(Nn, K, generator=g)
    Ad, Bd = ops.cast2d(A.to(cuda), torch.float32), ops.cast2d(B.to(cuda), torch.float32)
    ref = (A.double() @ B.double().t())
    scale = float(ref.abs().max())
    e_tf = float((ops.gemm_nt(Ad, Bd, backend=N.GEMM_TF32X3).cpu().double() - ref).abs().max())
    e_fp = float((ops.gemm_nt(Ad, Bd, backend=N.GEMM_SIMT).cpu().double() - ref).abs().max())
    assert e_tf <= 2.0 * e_fp + 1e-6 * scale, (e_tf, e_fp, scale)
    assert e_tf <= 2e-6 * scale * math.sqrt(K / 512), (e_tf, scale)
    auto = ops.gemm_nt(Ad, Bd)                      # AUTO picks the tensor-core path for fp32 operands
    assert torch.equal(auto, ops.gemm_nt(Ad, Bd, backend=N.GEMM_TF32X3))
    if M <= 4096:   # weight-gradient orientation on the same operands: [K, M]^T-style product reduced over M rows
        G = torch.randn(M, Nn, generator=g)
        Gd = ops.cast2d(G.to(cuda), torch.float32)
        ref_t = G.double().t() @ A.double()
        s_t = float(ref_t.abs().max())
        e_tf = float((ops.gemm_tn(Gd, Ad, backend=N.GEMM_TF32X3).cpu().double() - ref_t).abs().max())
        e_fp = float((ops.gemm_tn(Gd, Ad, backend=N.GEMM_SIMT).cpu().double() - ref_t).abs().max())
        assert e_tf <= 2.0 * e_fp + 1e-6 * s_t, (e_tf, e_fp, s_t)


def test_tcgen05_matches_simt_on_identical_bf16_inputs(cuda):
    g = torch.Generator().manual_seed(0)
    A, B = torch.randn(777, 320, generator=g).bfloat16().to(cuda), torch.randn(256, 320, generator=g).bfloat16().to(cuda)
    a = ops.gemm_nt(A, B, backend=N.GEMM_TCGEN05, out_dtype=torch.float32)
    b = ops.gemm_nt(A, B, backend=N.GEMM_SIMT, out_dtype=torch.float32)
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-3)  # same products, different fp32 summation order


def test_dropout_epilogue_simt_and_tcgen05(cuda):
    A = torch.ones(4096, 64, device=cuda)
    W = torch.eye(64, device=cuda)
    for backend, dt in ((N.GEMM_SIMT, torch.float32), (N.GEMM_TCGEN05, torch.bfloat16), (N.GEMM_TF32X3, torch.float32)):
        y1 = ops.gemm_nt(A.to(dt), W.to(dt), relu=True, dropout_p=0.5, seed=123, offset=7, backend=backend).float()
        y2 = ops.gemm_nt(A.to(dt), W.to(dt), relu=True, dropout_p=0.5, seed=123, offset=7, backend=backend).float()
        y3 = ops.gemm_nt(A.to(dt), W.to(dt), relu=True, dropout_p=0.5, seed=124, offset=7, backend=backend).float()
        assert torch.equal(y1, y2) and not torch.equal(y1, y3)             # counter-based, reproducible
        kept = (y1 > 0).float().mean().item()
        assert abs(kept - 0.5) < 0.01
        assert torch.equal(y1[y1 > 0], torch.full_like(y1[y1 > 0], 2.0))   # survivors scaled by 1/(1-p)
        gy = torch.ones_like(y1).to(dt)
        gz = ops.gate(gy, y1.to(dt), 2.0).float()
        assert torch.equal(gz, torch.where(y1 > 0, torch.full_like(y1, 2.0), torch.zeros_like(y1)))
    # both backends draw the same mask for the same (seed, offset)
    a = ops.gemm_nt(A, W, relu=True, dropout_p=0.3, seed=9, offset=1, backend=N.GEMM_SIMT) > 0
    b = ops.gemm_nt(A.bfloat16(), W.bfloat16(), relu=True, dropout_p=0.3, seed=9, offset=1, backend=N.GEMM_TCGEN05) > 0
    assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,Nn", [(300, 256), (77, 40), (5, 9), (0, 8)])
def test_add_act_equals_gemm_epilogue(cuda, dtype, M, Nn):
    """llp_add_act == the llp_gemm_nt epilogue (bias + addend + relu + dropout, identical Philox mask) applied to a
    pre-activation that is not a GEMM output — checked exactly through an identity GEMM, in the vector and scalar paths,
    for p = 0.5 (1-bit stream) and p = 0.3 (16-bit stream), on row-padded and strided views."""
    torch.manual_seed(3)
    wide = torch.randn(max(M, 1), 2 * Nn + 8, device=cuda).to(dtype)
    a, add = wide[:M, :Nn], wide[:M, Nn:2 * Nn]                       # strided views of one buffer (the [t | r] layout)
    bias = torch.randn(Nn, device=cuda)
    eye = torch.eye(Nn, device=cuda).to(dtype)
    backend = N.GEMM_SIMT
    for p, relu in ((0.0, False), (0.0, True), (0.5, True), (0.3, True)):
        y = ops.add_act(a, add, bias, relu=relu, dropout_p=p, seed=11, offset=5)
        assert y.shape == (M, Nn)
        if M == 0:
            continue
        ref = ops.gemm_nt(a.contiguous(), eye, bias=bias, addend=add.contiguous(), relu=relu, dropout_p=p, seed=11, offset=5,
                          backend=backend)
        assert torch.equal(y, ref), (p, relu)
    if M:
        y = ops.add_act(a, None, None, relu=True)
        assert torch.equal(y, torch.relu(a))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_colsum_cast_gate(cuda, dtype):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(3001, 70, generator=g).to(dtype)
    torch.testing.assert_close(ops.colsum(ops.cast2d(A.to(cuda), dtype)).cpu(), A.double().sum(0).float(), rtol=1e-5, atol=1e-3)
    T = ops.cast2d(A.to(cuda), torch.float32, transpose=True).cpu()
    assert torch.equal(T, A.float().t())
    assert torch.equal(ops.cast2d(A.float().to(cuda), torch.bfloat16).cpu(), A.float().bfloat16())


# ------------------------------------------------------------------------------------------------
# edge scoring
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,f,m", [(100, 256, 1000), (50, 36, 300), (10, 8, 0)])
def test_edge_hadamard_and_backward(cuda, dtype, n, f, m):
    g = torch.Generator().manual_seed(2)
    h = torch.randn(n, f, generator=g).to(dtype)
    u, v = torch.randint(0, n, (m,), generator=g), torch.randint(0, n, (m,), generator=g)
    hd = ops.cast2d(h.to(cuda), dtype).requires_grad_(True)
    z = ops.HadamardFn.apply(hd, u.to(cuda), v.to(cuda))
    ho = h.float().requires_grad_(True)
    zo = ho[u] * ho[v]
    tol = FP32 if dtype == torch.float32 else BF16
    torch.testing.assert_close(z.float().cpu(), zo.detach(), **tol)
    if m:
        gz = torch.randn(m, f, generator=g).to(dtype)
        z.backward(ops.cast2d(gz.to(cuda), dtype))
        zo.backward(gz.float())
        torch.testing.assert_close(hd.grad.float().cpu(), ho.grad, **(dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=3e-2, atol=0.25)))


@pytest.mark.parametrize("n,m,hubs", [(100, 1000, ()), (7, 0, ()), (1, 5, ()), (5000, 3, ()), (3000, 20000, (17, 300, 5000)),
                                      (235868, 131072, (40, 700)), (2200000, 4096, (2000,)), (9, 40000, ())])
def test_edge_plan_bit_exact(cuda, n, m, hubs):
    """llp_edge_plan (own counting sort) against the oracle's stable sort: row pointers and (edge, other endpoint) pairs
    index for index — empty batch, one node, rows of every length class (one thread / whole blocks / many slices),
    more scan tiles than one block handles at once."""
    g = torch.Generator().manual_seed(5)
    u, v = torch.randint(0, n, (m,), generator=g), torch.randint(0, n, (m,), generator=g)
    pos = 0
    for k, cnt in enumerate(hubs):   # node k + 1 gets `cnt` extra incidences, spread over both endpoint arrays
        (u if k % 2 == 0 else v)[pos:pos + cnt] = min(k + 1, n - 1)
        pos += cnt
    plan = ops.EdgePlan(u.to(cuda), v.to(cuda), n)
    rowptr, meta = O.edge_incidence_plan(u, v, n)
    assert torch.equal(plan.rowptr.cpu(), rowptr)
    assert torch.equal(plan.meta.cpu()[: 4 * m].view(-1, 2), meta)


def test_edge_hadamard_backward_bit_exact_fp32(cuda):
    """fp32 gather-reduce in plan order: the CUDA kernel adds fma(dz, h, acc) incidence by incidence, so it equals a
    float64-free restatement with the same order up to the fma's single rounding; against the oracle at 1e-6."""
    g = torch.Generator().manual_seed(6)
    n, f, m = 300, 64, 4000
    h = torch.randn(n, f, generator=g)
    u, v = torch.randint(0, n, (m,), generator=g), torch.randint(0, n, (m,), generator=g)
    u[:900] = 3   # a hub row (block kernel)
    dz = torch.randn(m, f, generator=g)
    plan = ops.EdgePlan(u.to(cuda), v.to(cuda), n)
    gh = ops.hadamard_bwd(h.to(cuda), plan, dz.to(cuda))
    want = O.hadamard_backward(h, u, v, dz)
    torch.testing.assert_close(gh.cpu(), want, rtol=1e-5, atol=1e-4)
    gh2 = ops.hadamard_bwd(h.to(cuda), ops.EdgePlan(u.to(cuda), v.to(cuda), n), dz.to(cuda))
    assert torch.equal(gh, gh2)   # bit-reproducible: no atomics on the data path


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_score_head_forward_backward(cuda, dtype):
    g = torch.Generator().manual_seed(3)
    M, H = 3000, 256
    y = (torch.randn(M, H, generator=g) * 0.3).to(dtype)
    w, b = torch.randn(1, H, generator=g) * 0.1, torch.randn(1, generator=g)
    yd = ops.cast2d(y.to(cuda), dtype).requires_grad_(True)
    wd, bd = w.to(cuda).requires_grad_(True), b.to(cuda).requires_grad_(True)
    p = ops.ScoreHeadFn.apply(yd, wd, bd)
    yo, wo, bo = y.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    po = torch.sigmoid(yo @ wo.t() + bo).squeeze(-1)
    torch.testing.assert_close(p.cpu(), po.detach(), rtol=1e-5, atol=1e-6)
    dp = torch.randn(M, generator=g)
    p.backward(dp.to(cuda))
    po.backward(dp)
    torch.testing.assert_close(wd.grad.cpu(), wo.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(bd.grad.cpu(), bo.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(yd.grad.float().cpu(), yo.grad, **(FP32 if dtype == torch.float32 else BF16))


# ------------------------------------------------------------------------------------------------
# dense negative sampling (train_teacher_gnn.py:50-51; main.py:81-82,206-207)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,e,want", [(60, 300, 150), (2708, 10556, 8976), (30, 400, 5), (30, 400, 40), (12, 100, 30), (500, 0, 64)])
def test_negative_sampling_dense_bit_exact(cuda, n, e, want):
    """shims.negative_sampling(method='dense') on the device (host random.sample -> side stream -> fused filter /
    compaction / de-linearisation; rounds 2 and 3 on the slow path) against the oracle's mask formulation: the same
    negatives index for index and the same advance of Python's random state, over seeds that do and do not need a
    second round (tiny requests on a half-full graph)."""
    import random

    from linkless_link_prediction_b200 import shims
    g = torch.Generator().manual_seed(3)
    ei = torch.randint(0, n, (2, e), generator=g)
    eid = ei.to(cuda)
    for seed in range(12):
        random.seed(seed); ref = O.negative_sampling_dense(ei, n, want); s_ref = random.getstate()
        random.seed(seed); got = shims.negative_sampling(eid, num_nodes=n, num_neg_samples=want, method="dense"); s_got = random.getstate()
        assert torch.equal(got.cpu(), ref), (n, e, want, seed)
        assert s_got == s_ref


def test_negative_filter_counts_beyond_capacity(cuda):
    """llp_negative_filter reports every kept candidate in `count` while writing only the first max_out."""
    lib = N.require_gpu()
    n = 50
    taken = torch.arange(0, n * (n - 1), 3, device=cuda)
    cand = torch.randperm(n * (n - 1), generator=torch.Generator().manual_seed(1))[:900].to(cuda)
    keep = ~torch.isin(cand, taken)
    want = cand[keep]
    for cap in (0, 10, int(want.numel()), 2000):
        kept = torch.full((max(cap, 1),), -1, dtype=torch.int64, device=cuda)
        edges = torch.full((2, max(cap, 1)), -1, dtype=torch.int64, device=cuda)
        cnt = torch.zeros(1, dtype=torch.int32, device=cuda)
        nb = lib.llp_negative_filter_workspace_bytes(cand.numel())
        ws = torch.empty(nb, dtype=torch.uint8, device=cuda)
        N.check(lib.llp_negative_filter(cand.data_ptr(), cand.numel(), taken.data_ptr(), taken.numel(), n, cap, kept.data_ptr(),
                                        edges.data_ptr(), cnt.data_ptr(), ws.data_ptr(), nb, N.stream_ptr()), "llp_negative_filter")
        assert int(cnt.item()) == want.numel()
        m = min(cap, want.numel())
        assert torch.equal(kept[:m], want[:m])
        r = want[:m] // (n - 1); c = want[:m] % (n - 1); c = c + (r <= c).long()
        if cap:
            assert torch.equal(edges.view(2, -1)[0, :m] if cap == 0 else edges.reshape(-1)[:m], r)
            assert torch.equal(edges.reshape(-1)[cap:cap + m], c)


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def test_bce_value_and_grad(cuda):
    g = torch.Generator().manual_seed(4)
    p = torch.rand(5000, generator=g).clamp(1e-6, 1 - 1e-6)
    p[0], p[1] = 0.0, 1.0  # log clamp at -100
    pd = p.to(cuda).requires_grad_(True)
    loss = ops.bce_loss(pd, 2000)
    po = p.clone().requires_grad_(True)
    lo = torch.nn.BCELoss()(po, torch.cat((torch.ones(2000), torch.zeros(3000))))
    torch.testing.assert_close(loss.cpu(), lo.detach(), **FP32)
    (loss * 3).backward()
    (lo * 3).backward()
    torch.testing.assert_close(pd.grad.cpu()[2:], po.grad[2:], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("rows,K,T", [(64, 12, 1.0), (1000, 36, 1.0), (7, 300, 2.0), (33, 1, 1.0)])
def test_llp_d_value_and_grad(cuda, rows, K, T):
    g = torch.Generator().manual_seed(5)
    s, t = torch.rand(rows, K, generator=g), torch.rand(rows, K, generator=g)
    sd = s.to(cuda).requires_grad_(True)
    loss = ops.kl_loss(sd, t.to(cuda), T)
    so = s.clone().requires_grad_(True)
    lo = O.kl_loss(so, t, T)
    torch.testing.assert_close(loss.cpu(), lo.detach(), rtol=1e-5, atol=1e-7)
    loss.backward()
    lo.backward()
    torch.testing.assert_close(sd.grad.cpu(), so.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("rows,K,margin", [(64, 12, 0.1), (500, 36, 0.01), (3, 100, 0.2), (10, 2, 0.05)])
def test_llp_r_value_and_grad(cuda, rows, K, margin):
    g = torch.Generator().manual_seed(6)
    s, t = torch.rand(rows, K, generator=g), torch.rand(rows, K, generator=g)
    t[0] = 0.5  # all teacher ties: every pair contributes the constant `margin`
    sd = s.to(cuda).requires_grad_(True)
    loss = ops.rank_loss(sd, t.to(cuda), margin)
    so = s.clone().requires_grad_(True)
    lo = O.llp_r_loss(so, t, margin)
    torch.testing.assert_close(loss.cpu(), lo.detach(), rtol=1e-5, atol=1e-7)
    loss.backward()
    lo.backward()
    torch.testing.assert_close(sd.grad.cpu(), so.grad, rtol=1e-4, atol=1e-8)
    assert torch.count_nonzero(sd.grad[0]) == 0


@pytest.mark.parametrize("rows,K,margin,wd,wr", [(64, 12, 0.1, 1.0, 1.0), (500, 36, 0.01, 0.5, 2.0), (3, 100, 0.2, 1.0, 0.0),
                                                  (2708, 12, 0.1, 1.0, 1.0), (10, 2, 0.05, 0.0, 1.0)])
def test_llp_fused_equals_separate_kernels(cuda, rows, K, margin, wd, wr):
    """LLP_D and LLP_R in one pass over the score rows (llp_kd_fused): both loss values equal to the separate
    kernels' (same arithmetic, same reduction tree), the combined gradient equal to wd * dLLP_D + wr * dLLP_R, and both against the oracle."""
    g = torch.Generator().manual_seed(rows + K)
    s = torch.sigmoid(torch.randn(rows, K, generator=g))
    t = torch.sigmoid(torch.randn(rows, K, generator=g))
    t[:, : K // 3] = t[:, :1]   # teacher ties: the constant-margin pairs of SURVEY.md Q6
    sd = s.to(cuda).requires_grad_(True)
    total, d, r = ops.kd_losses(sd, t.to(cuda), 1.0, margin, wd, wr)
    total.backward()
    s2 = s.to(cuda).requires_grad_(True)
    d2, r2 = ops.kl_loss(s2, t.to(cuda), 1.0), ops.rank_loss(s2, t.to(cuda), margin)
    (wd * d2 + wr * r2).backward()
    torch.testing.assert_close(d, d2.detach(), rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(r, r2.detach(), rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(total.detach(), wd * d2.detach() + wr * r2.detach(), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(sd.grad, s2.grad, rtol=1e-5, atol=1e-8)
    so = s.clone().requires_grad_(True)
    lo = wd * O.kl_loss(so, t, 1.0) + wr * O.llp_r_loss(so, t, margin)
    lo.backward()
    torch.testing.assert_close(total.detach().cpu(), lo.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(sd.grad.cpu(), so.grad, rtol=1e-4, atol=1e-7)


def test_golden_losses_from_reference(cuda, golden):
    g = golden["kl"]
    torch.testing.assert_close(ops.kl_loss(g["s"].to(cuda), g["t"].to(cuda), 1).cpu(), g["kl_T1"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ops.kl_loss(g["s"].to(cuda), g["t"].to(cuda), 2.0).cpu(), g["kl_T2"], rtol=1e-5, atol=1e-7)
    r = golden["llp_r"]
    torch.testing.assert_close(ops.rank_loss(r["s_r"].to(cuda), r["t_r"].to(cuda), r["margin"]).cpu(), r["loss"], rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------
# Hits@K, sampling, optimiser
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_pos,n_neg,Ks", [(1000, 5000, [10, 20, 30, 50]), (60084, 100000, [10, 50, 100]), (50, 30, [10, 50]),
                                            (10, 0, [1]), (777, 100, [100]), (5, 3000, [1, 2, 3])])
def test_hits_counts_bit_exact(cuda, n_pos, n_neg, Ks):
    from linkless_link_prediction_b200.shims import hits_counts
    g = torch.Generator().manual_seed(7)
    pos = torch.sigmoid(torch.randn(n_pos, generator=g) * 3)
    neg = torch.sigmoid(torch.randn(n_neg, generator=g) * 3)
    neg = (neg * 200).round() / 200  # heavy ties, also across the K-th value
    pos[: n_pos // 4] = (pos[: n_pos // 4] * 200).round() / 200
    counts, n = hits_counts(pos.to(cuda), neg.to(cuda), Ks)
    assert counts.tolist() == O.hits_counts(pos, neg, Ks)
    assert int(n) == n_pos
    if n_neg >= max(Ks):
        assert torch.equal(ops.topk_desc(neg.to(cuda), max(Ks)).cpu(), torch.topk(neg, max(Ks)).values)


@pytest.mark.parametrize("n_pos,n_neg,q", [(1000, 5000, 200), (60084, 100000, 0), (46329, 100000, 1000), (50, 30, 3),
                                           (1, 1, 0), (777, 1, 10), (0, 10, 0), (10, 0, 0), (300000, 1000000, 0)])
def test_auc_pairs_bit_exact(cuda, n_pos, n_neg, q):
    """N2: ROC-AUC as integer pair counts (llp_auc_pairs) == the oracle's, == sklearn within 1e-12 after the one
    division; heavy ties (q quantisation levels), the collab eval sizes, empty sides."""
    from linkless_link_prediction_b200.shims import roc_auc_score_device
    g = torch.Generator().manual_seed(9)
    pos = torch.sigmoid(torch.randn(n_pos, generator=g) * 3 + 0.3)
    neg = torch.sigmoid(torch.randn(n_neg, generator=g) * 3)
    if q:
        pos, neg = (pos * q).round() / q, (neg * q).round() / q
    pairs = ops.auc_pairs(pos.to(cuda), neg.to(cuda))
    assert tuple(pairs.tolist()) == O.auc_pairs(pos, neg)
    if n_pos and n_neg:
        assert roc_auc_score_device(pos.to(cuda), neg.to(cuda)) == O.roc_auc(pos, neg)
    else:
        with pytest.raises(ValueError):
            roc_auc_score_device(pos.to(cuda), neg.to(cuda))


def test_auc_signed_zero_and_negative_scores(cuda):
    pos = torch.tensor([-0.0, 0.0, -1.5, 2.0, 1e-30, -1e-30] * 40)
    neg = torch.tensor([0.0, -0.0, -1.0, 2.0, -7.0, 1e-30] * 33)
    assert tuple(ops.auc_pairs(pos.to(cuda), neg.to(cuda)).tolist()) == O.auc_pairs(pos, neg)


def test_topk_with_negative_scores_and_zeros(cuda):
    x = torch.tensor([-1.0, 0.0, -0.0, 3.5, -7.0, 3.5, 1e-30, -1e-30] * 50)
    assert torch.equal(ops.topk_desc(x.to(cuda), 130).cpu(), torch.topk(x, 130).values)


@pytest.mark.parametrize("n,e,B,L", [(200, 1500, 300, 2), (50, 100, 64, 6), (30, 0, 10, 3)])
def test_random_walk_bit_exact(cuda, n, e, B, L):
    from linkless_link_prediction_b200.shims import random_walk
    g = torch.Generator().manual_seed(8)
    row = torch.sort(torch.randint(0, n, (e,), generator=g)).values
    col = torch.randint(0, n, (e,), generator=g)
    start = torch.randint(0, n, (B,), generator=g)
    rand = torch.rand(B, L, generator=g)
    if e == 0:
        row, col = torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long)
    out = random_walk(row.to(cuda), col.to(cuda), start.to(cuda), L, coalesced=False, num_nodes=n, rand=rand.to(cuda))
    ref = O.random_walk_with_rand(O.walk_rowptr(row, n), col, start, rand)
    assert torch.equal(out.cpu(), ref)


def test_clip_adam_matches_torch(cuda):
    from linkless_link_prediction_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(64, 33), (64,), (7, 5), (1, 64), (1,)]
    ref = [torch.nn.Parameter(torch.randn(*s)) for s in shapes]
    dev = [torch.nn.Parameter(p.detach().clone().to(cuda)) for p in ref]
    opt_ref = torch.optim.Adam(ref, lr=0.01)
    opt = FusedAdam(dev, lr=0.01)
    for step in range(4):
        grads = [torch.randn(*s) * (5.0 if step % 2 == 0 else 0.01) for s in shapes]
        opt.zero_grad()
        for p, d, g in zip(ref, dev, grads):
            p.grad = g.clone()
            d.grad.copy_(g.to(cuda))
        torch.nn.utils.clip_grad_norm_(ref[:3], 1.0)
        torch.nn.utils.clip_grad_norm_(ref[3:], 1.0)
        opt_ref.step()
        opt.step(clip_groups=[dev[:3], dev[3:]], max_norm=1.0)
        for p, d in zip(ref, dev):
            torch.testing.assert_close(d.detach().cpu(), p.detach(), rtol=1e-5, atol=1e-6)
