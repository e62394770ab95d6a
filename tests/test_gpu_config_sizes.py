"""Parity at the sizes BASELINE.json names (VERDICT r01 "next round" item 1): the CUDA path against the CPU oracle on

* C1/C2  full Cora shape (2,708 nodes, 1,433-d features, 8,976 training messages): teacher epochs + evaluation, and the
         LLP student epoch (LLP_D = LLP_R = True_label = 1) distilled from that teacher;
* C3     Coauthor-Physics shape (34,493 nodes, 8,415-d features, SAGEConv_updated): one training step;
* C4     ogbl-collab shape (235,868 nodes, 2,358,104 messages, 128 -> 256 -> 256 -> 256): one training step (loss,
         embeddings, every gradient) and Hits@K counts at 60,084 positives / 100,000 negatives;
* C5     a power-law slice with hub rows of more than 10^5 edges: SpMM forward / transpose and one training step.

Tolerances (north_star): 1e-5 relative in fp32 mode, 2e-2 in bf16 mode for losses, logits and embeddings; integer
results (Hits@K counts) bit-exact on identical scores.  Parameter gradients are sums over up to 2.4 M terms with
cancellation AND they depend on the relu masks of every layer above: a pre-activation within round-off of zero takes
the other branch in any two fp32 implementations (measured at C4: ~50 of 60 M mask entries differ, each moving one row
of a weight gradient by 1/485 of its norm).  So gradients are bounded against the norm of the whole tensor, and in fp32
mode the bound follows the reference arithmetic's own distance from the truth: the oracle runs once more in fp64 and the
CUDA path (3xTF32 tensor-core GEMMs with per-k-block promotion) must be within 5x of the fp32 oracle's error + 2e-4 of
the norm (measured at C4, profiles/r02_fp32_accuracy.txt: 1e-6 .. 1.8e-4 against 4e-7 .. 4e-5 for the fp32 oracle and
5e-7 .. 1.5e-5 for the CUDA-core fp32 GEMMs; losses and embeddings agree to 1e-8 / 4e-7); bf16: 1e-1 of the norm (bf16
activations through three layers; measured 4.7e-2 at C4, 6.6e-2 at C3).
"""
import copy
import random

import numpy as np
import pytest
import torch

import linkless_link_prediction_b200 as L
from linkless_link_prediction_b200 import main as student
from linkless_link_prediction_b200 import ops, shims
from linkless_link_prediction_b200 import train_teacher_gnn as teacher
from linkless_link_prediction_b200.data import synthetic_dataset, undirected_graph
from oracle import llp_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float32: dict(rtol=1e-5, atol=2e-6), torch.bfloat16: dict(rtol=2e-2, atol=2e-2)}
LOSS_RTOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}
GRAD_REL_BF16 = 1e-1


@pytest.fixture(params=[torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def mode(request):
    ops.set_compute_dtype(request.param)
    yield request.param
    ops.set_compute_dtype(torch.bfloat16)


def seed_all(s):
    random.seed(s); np.random.seed(s); torch.manual_seed(s)


def _pair(cuda, f, H, layers, conv_o=None, conv_d=None, seed=0):
    """Oracle encoder + predictor and their device twins with identical parameters (dropout 0: the reference's dropout
    masks come from torch's RNG and cannot be reproduced by any other implementation)."""
    seed_all(seed)
    mo = O.SAGE("cfg", f, H, H, layers, 0.0, conv_o or O.SAGEConv)
    po = O.LinkPredictor("mlp", H, H, 1, 2, 0.0)
    md = L.SAGE("cfg", f, H, H, layers, 0.0, conv_d or L.SAGEConv); md.load_state_dict(mo.state_dict()); md.to(cuda)
    pd = L.LinkPredictor("mlp", H, H, 1, 2, 0.0); pd.load_state_dict(po.state_dict()); pd.to(cuda)
    return mo, po, md, pd


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def _assert_grads(named_o, named_d, mode, named_truth=None):
    report = {}
    for i, ((k, a), (_, b)) in enumerate(zip(named_o, named_d)):
        assert b.grad is not None, k
        if mode == torch.float32 and named_truth is not None:
            t = named_truth[i][1].grad
            e_cuda, e_ref = _rel(b.grad.cpu(), t), _rel(a.grad, t)
            report[k] = (e_cuda, e_ref)
            assert e_cuda <= 5.0 * e_ref + 2e-4, (k, e_cuda, e_ref)
        else:
            rel = _rel(b.grad.float().cpu(), a.grad)
            report[k] = rel
            assert rel < (GRAD_REL_BF16 if mode == torch.bfloat16 else 2e-3), (k, rel)
    return report


def _one_step_parity(cuda, mode, x, adj, pos, neg, mo, po, md, pd, check_h=True):
    """Forward + BCE + backward of one teacher step (train_teacher_gnn.py:37-61) on explicit edges: loss, embeddings and
    every parameter gradient, CUDA vs oracle."""
    mo.train(); po.train(); md.train(); pd.train()
    ho = mo(x, adj)
    edges = torch.cat((pos, neg), dim=-1)
    label = torch.cat((torch.ones(pos.size(1)), torch.zeros(neg.size(1))))
    lo = O.bce_loss(po(ho[edges[0]], ho[edges[1]]).squeeze(), label)
    lo.backward()
    truth = None
    if mode == torch.float32:   # the same step in fp64: what both fp32 implementations approximate
        m64, p64 = copy.deepcopy(mo).double(), copy.deepcopy(po).double()
        for q in list(m64.parameters()) + list(p64.parameters()):
            q.grad = None
        h64 = m64(x.double(), adj)
        O.bce_loss(p64(h64[edges[0]], h64[edges[1]]).squeeze(), label.double()).backward()
        truth = list(m64.named_parameters()) + list(p64.named_parameters())
        del h64
    hd = md(x.to(cuda), adj.to(cuda))
    ed = edges.to(cuda)
    ld = ops.bce_loss(pd.score(hd, ed[0].contiguous(), ed[1].contiguous()).reshape(-1), pos.size(1))
    ld.backward()
    assert float(ld.detach()) == pytest.approx(float(lo.detach()), rel=LOSS_RTOL[mode])
    if check_h:
        torch.testing.assert_close(hd.detach().float().cpu(), ho.detach(), **TOL[mode])
    _assert_grads(list(mo.named_parameters()) + list(po.named_parameters()),
                  list(md.named_parameters()) + list(pd.named_parameters()), mode, truth)
    return ho.detach(), hd.detach()


# ------------------------------------------------------------------------------------------------------------------
# C1 / C2: Cora shape
# ------------------------------------------------------------------------------------------------------------------
def test_c1_cora_full_shape_teacher_epochs_and_eval(cuda, mode):
    data, split = synthetic_dataset("cora", seed=0, scale=1.0)
    x, adj = data.x, data.adj_t
    assert tuple(x.shape) == (2708, 1433) and adj.size(1) == split["train"]["edge"].size(0)
    mo, po, md, pd = _pair(cuda, 1433, 256, 2)
    opt_o = torch.optim.Adam(list(mo.parameters()) + list(po.parameters()), lr=0.005)
    opt_d = L.FusedAdam(list(md.parameters()) + list(pd.parameters()), lr=0.005)
    dev_data = shims.Data(x=x, adj_t=adj).to(cuda)
    # epochs in the reference's RNG order (DataLoader shuffle + python random.sample negatives): same seed, same batches
    seed_all(5)
    lo = [O.teacher_train_epoch(mo, po, x, adj, split["train"]["edge"], opt_o, 65536, "sage", "cora") for _ in range(2)]
    seed_all(5)
    ld = [teacher.train(md, pd, dev_data, split, opt_d, 65536, "sage", "cora", "transductive") for _ in range(2)]
    assert ld[0] == pytest.approx(lo[0], rel=LOSS_RTOL[mode])          # identical weights: pure forward parity
    assert ld[1] == pytest.approx(lo[1], rel=1e-4 if mode == torch.float32 else 2e-2)   # after one clip + Adam update
    # evaluation from IDENTICAL weights (the device's, copied into the oracle): embeddings, Hits@K, AUC
    mo.load_state_dict({k: v.detach().cpu() for k, v in md.state_dict().items()})
    po.load_state_dict({k: v.detach().cpu() for k, v in pd.state_dict().items()})
    ro, ho = O.test_transductive(mo, po, x, adj, split, 65536, "sage", "cora")
    args = type("A", (), {"minibatch": False, "compute_auc": True})()
    rd, hd = teacher.test_transductive(md, pd, dev_data, split, L.Evaluator("ogbl-ddi"), 65536, "sage", "cora", args)
    torch.testing.assert_close(hd.float().cpu(), ho, **TOL[mode])
    n_pos = (split["valid"]["edge"].size(0), split["test"]["edge"].size(0))
    for K in (10, 20, 30, 50):
        for a, b, n in zip(rd[f"Hits@{K}"], ro[f"Hits@{K}"], n_pos):
            # scores agree to ~1e-6 (fp32): at most one positive may sit on the other side of the K-th negative
            slack = (1.0 if mode == torch.float32 else 0.08 * n) / n
            assert abs(a - b) <= slack + 1e-12, (K, a, b)
    for a, b in zip(rd["AUC"], ro["AUC"]):
        assert a == pytest.approx(b, abs=1e-4 if mode == torch.float32 else 2e-2)
    # integer part: Hits@K counts on IDENTICAL scores are bit-exact
    md.eval(); pd.eval()
    with torch.no_grad():
        e, en = split["valid"]["edge"].to(cuda), split["valid"]["edge_neg"].to(cuda)
        sp = pd.score(hd, e[:, 0].contiguous(), e[:, 1].contiguous()).reshape(-1)
        sn = pd.score(hd, en[:, 0].contiguous(), en[:, 1].contiguous()).reshape(-1)
    counts, _ = shims.hits_counts(sp, sn, [10, 20, 30, 50])
    assert counts.tolist() == O.hits_counts(sp.float().cpu(), sn.float().cpu(), [10, 20, 30, 50])


def test_c2_cora_full_shape_student_epoch(cuda, mode):
    """LLP student (MLP 1433 -> 256 -> 256, LLP_D = LLP_R = True_label = 1; README.md:26) distilled from a SAGE teacher's
    embeddings: one epoch of main.train() in the reference's RNG order vs the oracle's restatement of that loop."""
    data, split = synthetic_dataset("cora", seed=0, scale=1.0)
    x, adj = data.x, data.adj_t
    seed_all(1)
    t_model = O.SAGE("cora", 1433, 256, 256, 2, 0.0).eval()
    t_pred_o = O.LinkPredictor("mlp", 256, 256, 1, 2, 0.0)
    with torch.no_grad():
        t_h = t_model(x, adj)
    so = O.MLP(2, 1433, 256, 256, 0.0)
    po = O.LinkPredictor("mlp", 256, 256, 1, 2, 0.0)
    sd = L.MLP(2, 1433, 256, 256, 0.0); sd.load_state_dict(so.state_dict()); sd.to(cuda)
    pd = L.LinkPredictor("mlp", 256, 256, 1, 2, 0.0); pd.load_state_dict(po.state_dict()); pd.to(cuda)
    t_pred_d = L.LinkPredictor("mlp", 256, 256, 1, 2, 0.0); t_pred_d.load_state_dict(t_pred_o.state_dict()); t_pred_d.to(cuda)
    for p in list(t_pred_o.parameters()) + list(t_pred_d.parameters()):
        p.requires_grad = False
    n_train = split["train"]["edge"].size(0)
    args = type("A", (), dict(transductive="transductive", link_batch_size=65536,
                              node_batch_size=int(x.size(0) / (n_train / 65536)), LLP_R=1.0, LLP_D=1.0, True_label=1.0,
                              KD_RM=0.0, KD_LM=0.0, margin=0.1, rw_step=3, ps_method="nb", ns_rate=1, hops=2,
                              datasets="cora"))()
    opt_o = torch.optim.Adam(list(so.parameters()) + list(po.parameters()), lr=0.005)
    opt_d = L.FusedAdam(list(sd.parameters()) + list(pd.parameters()), lr=0.005)
    dev_data = shims.Data(x=x, adj_t=adj).to(cuda)
    seed_all(7)
    lo = [O.student_train_epoch(so, po, t_h, t_pred_o, x, adj, split["train"]["edge"], opt_o, args, "cora") for _ in range(2)]
    shims.draw_rand_on_host(True)   # walks / random contexts from torch's CPU generator: the stream the oracle consumes
    try:
        seed_all(7)
        ld = [student.train(sd, pd, t_h.to(cuda), t_pred_d, dev_data, split, opt_d, args, cuda) for _ in range(2)]
    finally:
        shims.draw_rand_on_host(False)
    assert ld[0] == pytest.approx(lo[0], rel=LOSS_RTOL[mode])
    assert ld[1] == pytest.approx(lo[1], rel=1e-4 if mode == torch.float32 else 2e-2)


def test_c2_student_minibatch_matches_oracle_step(cuda, mode):
    """train_minibatch (main.py:52-144) encodes only the rows a step touches.  Its loss is checked against the ORACLE's
    student step on the same samples / edges (not against this package's own full-batch path)."""
    n, f, H = 2708, 1433, 256
    data, split = synthetic_dataset("cora", seed=0, scale=1.0)
    x, adj = data.x, data.adj_t
    seed_all(2)
    t_h = torch.randn(n, H) * 0.3
    so = O.MLP(2, f, H, H, 0.0); po = O.LinkPredictor("mlp", H, H, 1, 2, 0.0); to = O.LinkPredictor("mlp", H, H, 1, 2, 0.0)
    sd = L.MLP(2, f, H, H, 0.0); sd.load_state_dict(so.state_dict()); sd.to(cuda)
    pd = L.LinkPredictor("mlp", H, H, 1, 2, 0.0); pd.load_state_dict(po.state_dict()); pd.to(cuda)
    td = L.LinkPredictor("mlp", H, H, 1, 2, 0.0); td.load_state_dict(to.state_dict()); td.to(cuda)
    n_train = split["train"]["edge"].size(0)
    args = type("A", (), dict(transductive="transductive", link_batch_size=65536,
                              node_batch_size=int(n / (n_train / 65536)), LLP_R=1.0, LLP_D=1.0, True_label=1.0, KD_RM=0.0,
                              KD_LM=0.0, margin=0.1, rw_step=3, ps_method="nb", ns_rate=1, hops=2, datasets="cora"))()
    opt_o = torch.optim.Adam(list(so.parameters()) + list(po.parameters()), lr=0.005)
    opt_d = L.FusedAdam(list(sd.parameters()) + list(pd.parameters()), lr=0.005)
    seed_all(9)
    lo = O.student_train_epoch(so, po, t_h, to, x, adj, split["train"]["edge"], opt_o, args, "cora")
    shims.draw_rand_on_host(True)
    try:
        seed_all(9)
        # reference order inside train_minibatch: negative edges are drawn BEFORE the context samples (main.py:81-91),
        # the full-batch loop draws them after (:180,:205); python's random and torch's generator are separate streams,
        # so the same seeds still give the same batches
        ld = student.train_minibatch(sd, pd, t_h.to(cuda), td, shims.Data(x=x, adj_t=adj).to(cuda), split, opt_d, args, cuda)
    finally:
        shims.draw_rand_on_host(False)
    assert ld == pytest.approx(lo, rel=LOSS_RTOL[mode])


# ------------------------------------------------------------------------------------------------------------------
# C3: Coauthor-Physics shape, SAGEConv_updated (train_teacher_gnn.py:376-379)
# ------------------------------------------------------------------------------------------------------------------
def test_c3_physics_shape_training_step(cuda, mode):
    data, split = synthetic_dataset("coauthor-physics", seed=0, scale=1.0)
    x, adj = data.x, data.adj_t
    assert tuple(x.shape) == (34493, 8415)
    mo, po, md, pd = _pair(cuda, 8415, 256, 2, O.SAGEConvUpdated, L.SAGEConv_updated)
    g = torch.Generator().manual_seed(3)
    pos_all = split["train"]["edge"]
    pos = pos_all[torch.randperm(pos_all.size(0), generator=g)[:65536]].t().contiguous()
    neg = torch.randint(0, x.size(0), pos.size(), generator=g)
    _one_step_parity(cuda, mode, x, adj, pos, neg, mo, po, md, pd)


# ------------------------------------------------------------------------------------------------------------------
# C4: ogbl-collab shape
# ------------------------------------------------------------------------------------------------------------------
def test_c4_collab_shape_training_step_and_hits(cuda, mode):
    data, split = synthetic_dataset("collab", seed=0, scale=1.0)
    x, adj = data.x, data.adj_t
    assert x.size(0) == 235868 and adj.size(1) == 2358104 and x.size(1) == 128
    mo, po, md, pd = _pair(cuda, 128, 256, 3)
    g = torch.Generator().manual_seed(4)
    pos_all = split["train"]["edge"]
    pos = pos_all[torch.randint(0, pos_all.size(0), (65536,), generator=g)].t().contiguous()
    neg = torch.randint(0, x.size(0), pos.size(), generator=g)      # the collab branch, train_teacher_gnn.py:52-54
    _, hd = _one_step_parity(cuda, mode, x, adj, pos, neg, mo, po, md, pd)
    # Hits@{10,50,100} counts at 60,084 positives / 100,000 negatives: bit-exact on identical scores, and the scoring
    # itself against the oracle's predictor on the same embeddings
    assert split["valid"]["edge"].size(0) == 60084 and split["valid"]["edge_neg"].size(0) == 100000
    md.eval(); pd.eval(); po.eval()
    with torch.no_grad():
        e, en = split["valid"]["edge"].to(cuda), split["valid"]["edge_neg"].to(cuda)
        sp = pd.score(hd, e[:, 0].contiguous(), e[:, 1].contiguous()).reshape(-1)
        sn = pd.score(hd, en[:, 0].contiguous(), en[:, 1].contiguous()).reshape(-1)
        hc = hd.float().cpu()
        sp_o = O.score_edges(po, hc, split["valid"]["edge"], 65536)
    torch.testing.assert_close(sp.float().cpu(), sp_o, **TOL[mode])
    counts, n_pos = shims.hits_counts(sp, sn, [10, 50, 100])
    assert int(n_pos) == 60084
    assert counts.tolist() == O.hits_counts(sp.float().cpu(), sn.float().cpu(), [10, 50, 100])
    pairs = ops.auc_pairs(sp, sn).tolist()
    assert tuple(pairs) == O.auc_pairs(sp.float().cpu(), sn.float().cpu())


# ------------------------------------------------------------------------------------------------------------------
# C5: power-law slice with real hub rows (> 10^5 edges)
# ------------------------------------------------------------------------------------------------------------------
def _hub_graph(n=200_000, pairs=1_500_000, hubs=((7, 150_000), (11, 120_000)), seed=1):
    """Chung-Lu power-law graph + explicit hub nodes linked (both directions) to `deg` distinct random nodes."""
    ei = undirected_graph(n, pairs, seed, True)
    g = torch.Generator().manual_seed(seed + 100)
    extra = []
    for node, deg in hubs:
        other = torch.randperm(n, generator=g)[:deg]
        other = other[other != node]
        me = torch.full_like(other, node)
        extra += [torch.stack([other, me]), torch.stack([me, other])]
    ei = torch.cat([ei] + extra, dim=1)
    return ei[:, torch.randperm(ei.size(1), generator=g)].contiguous()   # hub edges scattered through the edge list


def _oracle_spmm64(ei, x, n, transpose, inv_deg=None, chunk=250_000):
    """fp64 gather-scatter (PyG's path) in edge chunks so the [E, F] message tensor never exists."""
    src, dst = (ei[1], ei[0]) if transpose else (ei[0], ei[1])
    out = torch.zeros(n, x.size(1), dtype=torch.float64)
    for s in range(0, ei.size(1), chunk):
        m = x[src[s:s + chunk]].double()
        if transpose:
            m = m * inv_deg[src[s:s + chunk]].double().unsqueeze(1)
        out.index_add_(0, dst[s:s + chunk], m)
    return out


def test_c5_power_law_hub_rows_spmm_and_step(cuda, mode):
    n, F = 200_000, 256
    ei = _hub_graph(n)
    deg = torch.bincount(ei[1], minlength=n)
    assert int(deg.max()) >= 100_000 and int((deg > 100_000).sum()) >= 2
    g = torch.Generator().manual_seed(2)
    x = torch.randn(n, F, generator=g)
    gy = torch.randn(n, F, generator=g)
    if mode == torch.bfloat16:   # the kernel's inputs ARE bf16 in this mode: give the oracle the same rounded values
        x, gy = x.bfloat16().float(), gy.bfloat16().float()
    graph = ops.Graph(ei.to(cuda), n)
    assert graph.hubs[1] >= 2 and graph.t_hubs[1] >= 2
    inv_deg = 1.0 / deg.clamp(min=1).float()
    fwd = graph.spmm(ops.to_compute(x.to(cuda))).float().cpu()
    ref = (_oracle_spmm64(ei, x, n, False) / deg.clamp(min=1).double().unsqueeze(1)).float()
    tol = dict(rtol=1e-5, atol=1e-6) if mode == torch.float32 else dict(rtol=1e-2, atol=2e-3)
    torch.testing.assert_close(fwd, ref, **tol)
    bwd = graph.spmm(ops.to_compute(gy.to(cuda)), transpose=True).float().cpu()
    ref_t = _oracle_spmm64(ei, gy, n, True, inv_deg).float()
    # the transpose of a hub row sums 150,000 signed terms of size ~|g| / deg(d): single elements cancel to ~0 while the
    # row's scale is ~50, so the error is bounded against each ROW's largest magnitude
    row_err = (bwd - ref_t).abs().amax(dim=1) / ref_t.abs().amax(dim=1).clamp_min(1e-3)
    assert float(row_err.max()) < (1e-5 if mode == torch.float32 else 1e-2), (float(row_err.max()), int(row_err.argmax()))
    assert float(row_err[7]) < (2e-6 if mode == torch.float32 else 1e-2) and float(row_err[11]) < (2e-6 if mode == torch.float32 else 1e-2)
    # one full training step on the same graph (hub rows through the encoder, its backward and the scorer's gather)
    f_in, H = 64, 64
    xs = x[:, :f_in].contiguous()
    mo, po, md, pd = _pair(cuda, f_in, H, 2)
    pos = ei[:, torch.randint(0, ei.size(1), (65536,), generator=g)].contiguous()
    neg = torch.randint(0, n, pos.size(), generator=g)
    _one_step_parity(cuda, mode, xs, ei, pos, neg, mo, po, md, pd)
