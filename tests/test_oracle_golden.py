"""The oracle against the fixtures produced by the REFERENCE'S OWN functions (tests/golden/make_golden.py),
plus first-principles known-answer tests for the third-party semantics it restates."""
import random

import numpy as np
import pytest
import torch

from oracle import llp_oracle as O

RTOL = 1e-5  # fp32 tolerance stated by BASELINE.json north_star


def test_models_match_reference(golden):
    g = golden["models"]
    mlp = O.MLP(3, 24, 32, 16, 0.5).eval(); mlp.load_state_dict(g["mlp_sd"])
    sage = O.SAGE("cora", 24, 32, 16, 3, 0.5, O.SAGEConv).eval(); sage.load_state_dict(g["sage_sd"])
    sage_u = O.SAGE("p", 24, 32, 16, 2, 0.5, O.SAGEConvUpdated).eval(); sage_u.load_state_dict(g["sage_u_sd"])
    pred = O.LinkPredictor("mlp", 16, 32, 1, 3, 0.5).eval(); pred.load_state_dict(g["pred_sd"])
    pred_in = O.LinkPredictor("inner", 16, 32, 1, 2, 0.5).eval(); pred_in.load_state_dict(g["pred_in_sd"])
    with torch.no_grad():
        torch.testing.assert_close(mlp(g["x"]), g["mlp_out"], rtol=RTOL, atol=1e-6)
        torch.testing.assert_close(sage(g["x"], g["edge_index"]), g["sage_out"], rtol=RTOL, atol=1e-6)
        torch.testing.assert_close(sage_u(g["x"], g["edge_index"]), g["sage_u_out"], rtol=RTOL, atol=1e-6)
        torch.testing.assert_close(pred(g["xi"], g["xj"]), g["pred_out"], rtol=RTOL, atol=1e-6)
        torch.testing.assert_close(pred(g["x3i"], g["x3j"]), g["pred_out3"], rtol=RTOL, atol=1e-6)
        torch.testing.assert_close(pred_in(g["xi"], g["xj"]), g["pred_in_out"], rtol=RTOL, atol=1e-6)


def test_kl_and_rank_losses_match_reference(golden):
    g = golden["kl"]
    torch.testing.assert_close(O.kl_loss(g["s"], g["t"], 1), g["kl_T1"], rtol=RTOL, atol=1e-7)
    torch.testing.assert_close(O.kl_loss(g["s"], g["t"], 2.0), g["kl_T2"], rtol=RTOL, atol=1e-7)
    torch.testing.assert_close(O.cosine_loss(g["s"], g["t"]), g["cos"], rtol=RTOL, atol=1e-7)
    r = golden["llp_r"]
    torch.testing.assert_close(O.llp_r_loss(r["s_r"], r["t_r"], r["margin"]), r["loss"], rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("tag", ["teacher_fullbatch", "teacher_minibatch"])
def test_teacher_epoch_matches_reference(golden, tag):
    g = golden[tag]
    x, split, H = g["x"], g["split"], g["H"]
    adj_t = split["train"]["edge"].t().contiguous()
    model = O.SAGE("cora", x.size(1), H, H, 2, 0.0, O.SAGEConv)
    predictor = O.LinkPredictor("mlp", H, H, 1, 2, 0.0)
    model.load_state_dict(g["sd0"]["gnn"]); predictor.load_state_dict(g["sd0"]["predictor"])
    opt = torch.optim.Adam(list(model.parameters()) + list(predictor.parameters()), lr=g["lr"])
    random.seed(g["seed_train"]); np.random.seed(g["seed_train"]); torch.manual_seed(g["seed_train"])
    losses = [O.teacher_train_epoch(model, predictor, x, adj_t, split["train"]["edge"], opt, g["batch_size"]) for _ in range(2)]
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)
    for k, v in model.state_dict().items():
        torch.testing.assert_close(v, g["sd1"]["gnn"][k], rtol=1e-4, atol=1e-6)
    results, h = O.test_transductive(model, predictor, x, adj_t, split, g["batch_size"])
    torch.testing.assert_close(h, g["h"], rtol=1e-4, atol=1e-6)
    for K in (10, 20, 30, 50):
        assert results[f"Hits@{K}"] == pytest.approx(g["results"][f"Hits@{K}"], abs=1e-12)
    # the golden AUC is sklearn's roc_auc_score called by the reference's own test_transductive (:153)
    assert results["AUC"] == pytest.approx(g["results"]["AUC"], abs=1e-4)  # one flipped pair of 263 x 263 = 1.4e-5 (CPU reductions are not run-to-run deterministic)


def test_student_step_matches_reference(golden):
    g = golden["student"]
    a = g["args"]
    x, split, H, t_h = g["x"], g["split"], g["H"], g["t_h"]
    model = O.MLP(2, x.size(1), H, H, 0.0); model.load_state_dict(g["sd0"]["mlp"])
    pred = O.LinkPredictor("mlp", H, H, 1, 2, 0.0); pred.load_state_dict(g["sd0"]["predictor"])
    t_pred = O.LinkPredictor("mlp", H, H, 1, 2, 0.0); t_pred.load_state_dict(g["teacher_pred_sd"])
    opt = torch.optim.Adam(list(model.parameters()) + list(pred.parameters()), lr=g["lr"])
    random.seed(g["seed_train"]); np.random.seed(g["seed_train"]); torch.manual_seed(g["seed_train"])
    # the oracle's epoch loop (main.py:147-236 in the reference's RNG order) against the reference's own train()
    args = type("A", (), dict(a))()
    adj_t = split["train"]["edge"].t()
    losses = [O.student_train_epoch(model, pred, t_h, t_pred, x, adj_t, split["train"]["edge"], opt, args, "cora")
              for _ in range(2)]
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)


# ---- first-principles known answers for the third-party semantics (parity unpinned by the reference) ----
def test_mean_aggregate_kat():
    # path 0->1->2, star into 3, isolated 4, duplicate edge 0->1, self loop 2->2
    ei = torch.tensor([[0, 1, 0, 1, 2, 0, 2], [1, 2, 3, 3, 3, 1, 2]])
    x = torch.tensor([[1., 10.], [2., 20.], [4., 40.], [8., 80.], [16., 160.]])
    out = O.mean_aggregate(x, ei)
    exp = torch.tensor([[0., 0.], [1., 10.], [(2 + 4) / 2, (20 + 40) / 2], [(1 + 2 + 4) / 3, (10 + 20 + 40) / 3], [0., 0.]])
    torch.testing.assert_close(out, exp)
    rowptr, col, perm = O.csr_build(ei, 5)
    assert rowptr.tolist() == [0, 0, 2, 4, 7, 7]
    assert col.tolist() == [0, 0, 1, 2, 0, 1, 2]          # stable in edge order
    assert perm.tolist() == [0, 5, 1, 6, 2, 3, 4]
    torch.testing.assert_close(O.spmm_csr(rowptr, col, x, True), exp)


def test_mean_aggregate_matches_scipy_and_networkx():
    """Independent pins of the aggregation the third-party half of the path performs (PyG SAGEConv(aggr='mean') ->
    torch_scatter.scatter(reduce='mean'), absent here): the row-normalised adjacency product computed by scipy.sparse
    (multi-edges counted with their multiplicity, isolated nodes -> zero rows) and a per-node neighbour average over a
    networkx MultiDiGraph.  Forward and the autograd transpose (A~^T g)."""
    import networkx as nx
    import scipy.sparse as sp
    g = torch.Generator().manual_seed(7)
    n, f, e = 60, 5, 400
    ei = torch.randint(0, n - 6, (2, e), generator=g)          # multi-edges and self loops occur; the last 6 nodes are isolated
    x = torch.randn(n, f, generator=g, dtype=torch.float64)
    A = sp.coo_matrix((np.ones(e), (ei[1].numpy(), ei[0].numpy())), shape=(n, n)).tocsr()   # A[dst, src] = multiplicity
    deg = np.asarray(A.sum(1)).ravel()
    An = sp.diags(1.0 / np.maximum(deg, 1.0)) @ A
    ours = O.mean_aggregate(x, ei, n)
    np.testing.assert_allclose(ours.numpy(), An @ x.numpy(), rtol=1e-12, atol=1e-12)
    assert torch.equal(ours[n - 6:], torch.zeros(6, f, dtype=torch.float64))
    G = nx.MultiDiGraph(); G.add_nodes_from(range(n)); G.add_edges_from(zip(ei[0].tolist(), ei[1].tolist()))
    for v in (0, 3, 17, n - 1):
        preds = [u for u, _ in G.in_edges(v)]                 # with multiplicity
        want = x[preds].mean(0) if preds else torch.zeros(f, dtype=torch.float64)
        torch.testing.assert_close(ours[v], want, rtol=1e-12, atol=1e-12)
    # the transpose autograd applies: d/dx sum(w * mean_aggregate(x)) = A~^T w
    xg = x.clone().requires_grad_(True)
    w = torch.randn(n, f, generator=g, dtype=torch.float64)
    (O.mean_aggregate(xg, ei, n) * w).sum().backward()
    np.testing.assert_allclose(xg.grad.numpy(), An.T @ w.numpy(), rtol=1e-12, atol=1e-12)
    # and the CSR form the CUDA kernels are compared against
    rp, col, _ = O.csr_build(ei, n, "dst")
    np.testing.assert_allclose(O.spmm_csr(rp, col, x, mean=True).numpy(), An @ x.numpy(), rtol=1e-12, atol=1e-12)


def test_edge_incidence_plan_kat_and_autograd():
    """The incidence plan (backward of the gathers h[u] * h[v], train_teacher_gnn.py:58 / models.py:140): a hand-checked
    case, and its gather-reduce against what torch autograd (index_put_ with accumulation) produces for the same op."""
    u, v = torch.tensor([2, 0, 2, 3]), torch.tensor([0, 2, 2, 0])
    rowptr, meta = O.edge_incidence_plan(u, v, 5)
    assert rowptr.tolist() == [0, 3, 3, 7, 8, 8]
    # node 0: incidences 1 (u[1]), 4 (v[0]), 7 (v[3]); node 2: 0, 2, 5 (v[1]), 6 (v[2]); node 3: 3
    assert meta.tolist() == [[1, 2], [0, 2], [3, 3], [0, 0], [2, 2], [1, 0], [2, 2], [3, 0]]
    g = torch.Generator().manual_seed(11)
    n, f, m = 40, 6, 300
    h = torch.randn(n, f, generator=g, dtype=torch.float64).requires_grad_(True)
    uu, vv = torch.randint(0, n - 4, (m,), generator=g), torch.randint(0, n - 4, (m,), generator=g)
    dz = torch.randn(m, f, generator=g, dtype=torch.float64)
    (h[uu] * h[vv]).backward(dz)
    torch.testing.assert_close(O.hadamard_backward(h.detach(), uu, vv, dz).double(), h.grad, rtol=1e-6, atol=1e-6)
    rp, mt = O.edge_incidence_plan(uu, vv, n)
    assert int(rp[-1]) == 2 * m and torch.equal(rp[n - 4:], torch.full((5,), 2 * m, dtype=torch.int32))
    e0, e1 = O.edge_incidence_plan(torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long), 3)
    assert e0.tolist() == [0, 0, 0, 0] and e1.shape == (0, 2)


def test_sageconv_isolated_node_bias_quirk():
    # SAGEConv keeps b_l on isolated nodes, SAGEConv_updated aggregates it away
    torch.manual_seed(0)
    ei = torch.tensor([[0], [1]])
    x = torch.randn(3, 4)
    a, b = O.SAGEConv(4, 5), O.SAGEConvUpdated(4, 5)
    b.load_state_dict(a.state_dict())
    ya, yb = a(x, ei), b(x, ei)
    torch.testing.assert_close(ya[1], yb[1], rtol=1e-5, atol=1e-6)            # node with a neighbour: identical
    torch.testing.assert_close(ya[2] - yb[2], a.lin_l.bias.detach(), rtol=1e-5, atol=1e-6)


def test_hits_at_k_kat():
    pos = torch.tensor([0.9, 0.5, 0.5, 0.1])
    neg = torch.tensor([0.5, 0.5, 0.3, 0.2])
    assert O.hits_at_k(pos, neg, 1) == 0.25     # thr 0.5, strict '>' : only 0.9
    assert O.hits_at_k(pos, neg, 3) == 0.75     # thr 0.3
    assert O.hits_at_k(pos, neg, 5) == 1.0      # fewer negatives than K
    assert O.hits_counts(pos, neg, [1, 3, 5]) == [1, 3, 4]


def test_roc_auc_matches_sklearn_and_kat():
    """N2: the integer pair-count form of ROC-AUC against sklearn's trapezoid (the reference's call,
    train_teacher_gnn.py:147-153), including heavy ties, -0.0 == +0.0 and the one-class error."""
    from sklearn.metrics import roc_auc_score
    pos = torch.tensor([0.9, 0.5, 0.5, 0.1])
    neg = torch.tensor([0.5, 0.5, 0.3, 0.2])
    # pairs neg<pos: 0.9 -> 4, 0.5 -> 2 each, 0.1 -> 0 ; ties: 0.5 x 0.5 -> 2 each
    assert O.auc_pairs(pos, neg) == (8, 4)
    assert O.roc_auc(pos, neg) == (8 + 2) / 16
    assert O.auc_pairs(torch.tensor([-0.0, 0.0]), torch.tensor([0.0, -0.0, -1.0])) == (2, 4)
    g = torch.Generator().manual_seed(11)
    for n_pos, n_neg, q in ((1, 1, 0), (50, 3000, 0), (2000, 1500, 100), (4000, 4000, 7), (263, 263, 0)):
        p = torch.sigmoid(torch.randn(n_pos, generator=g) * 3 + 0.5)
        n = torch.sigmoid(torch.randn(n_neg, generator=g) * 3)
        if q:
            p, n = (p * q).round() / q, (n * q).round() / q
        y = np.concatenate((np.ones(n_pos), np.zeros(n_neg)))
        ref = roc_auc_score(y, torch.cat((p, n)).numpy())
        assert abs(O.roc_auc(p, n) - ref) <= 1e-12, (n_pos, n_neg, q)
    with pytest.raises(ValueError):
        O.roc_auc(torch.zeros(0), neg)


def test_llp_r_margin_constant_term():
    # teacher ties (|ti-tj| <= m) give y=0 -> each such pair contributes exactly `margin`
    s = torch.tensor([[0.2, 0.9, 0.4]])
    t = torch.tensor([[0.5, 0.5, 0.5]])
    assert O.llp_r_loss(s, t, 0.1).item() == pytest.approx(0.1)


def test_random_walk_kat():
    row = torch.tensor([0, 0, 1, 2, 2, 2])
    col = torch.tensor([1, 2, 2, 0, 1, 3])
    rowptr = O.walk_rowptr(row, 4)
    assert rowptr.tolist() == [0, 2, 3, 6, 6]
    rand = torch.tensor([[0.0, 0.99, 0.5], [0.7, 0.4, 0.0]])
    out = O.random_walk_with_rand(rowptr, col, torch.tensor([0, 2]), rand)
    # walk 0: 0 -(0.0*2=0)-> 1 -(0.99*1=0)-> 2 -(0.5*3=1)-> 1 ; walk 1: 2 -(0.7*3=2)-> 3 (sink) stays, stays
    assert out.tolist() == [[0, 1, 2, 1], [2, 3, 3, 3]]


def test_negative_sampling_never_returns_edges_or_self_loops():
    ei = O.synthetic_undirected_graph(60, 200, seed=1)
    random.seed(3)
    neg = O.negative_sampling_dense(ei, 60, 150)
    assert neg.shape == (2, 150)
    assert (neg[0] != neg[1]).all()
    have = set((ei[0] * 60 + ei[1]).tolist())
    assert not (set((neg[0] * 60 + neg[1]).tolist()) & have)
