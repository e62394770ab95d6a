"""End-to-end parity on the GPU: the drop-in modules and step functions against (a) fixtures produced by the
reference's own code (tests/golden/reference_golden.pt) and (b) the CPU oracle on the same seeded inputs."""
import random

import numpy as np
import pytest
import torch

import linkless_link_prediction_b200 as L
from linkless_link_prediction_b200 import main as student
from linkless_link_prediction_b200 import ops, shims
from linkless_link_prediction_b200 import train_teacher_gnn as teacher
from oracle import llp_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float32: dict(rtol=1e-5, atol=2e-6), torch.bfloat16: dict(rtol=2e-2, atol=2e-2)}


@pytest.fixture(params=[torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def mode(request):
    ops.set_compute_dtype(request.param)
    yield request.param
    ops.set_compute_dtype(torch.bfloat16)


def seed_all(s):
    random.seed(s); np.random.seed(s); torch.manual_seed(s)


def test_models_forward_match_reference_golden(cuda, golden, mode):
    g = golden["models"]
    x, ei = g["x"].to(cuda), g["edge_index"].to(cuda)
    mlp = L.MLP(3, 24, 32, 16, 0.5).eval(); mlp.load_state_dict(g["mlp_sd"]); mlp.to(cuda)
    sage = L.SAGE("cora", 24, 32, 16, 3, 0.5, L.SAGEConv).eval(); sage.load_state_dict(g["sage_sd"]); sage.to(cuda)
    sage_u = L.SAGE("p", 24, 32, 16, 2, 0.5, L.SAGEConv_updated).eval(); sage_u.load_state_dict(g["sage_u_sd"]); sage_u.to(cuda)
    pred = L.LinkPredictor("mlp", 16, 32, 1, 3, 0.5).eval(); pred.load_state_dict(g["pred_sd"]); pred.to(cuda)
    pred_in = L.LinkPredictor("inner", 16, 32, 1, 2, 0.5).eval(); pred_in.load_state_dict(g["pred_in_sd"]); pred_in.to(cuda)
    tol = TOL[mode]
    with torch.no_grad():
        torch.testing.assert_close(mlp(x).float().cpu(), g["mlp_out"], **tol)
        torch.testing.assert_close(sage(x, ei).float().cpu(), g["sage_out"], **tol)
        torch.testing.assert_close(sage_u(x, ei).float().cpu(), g["sage_u_out"], **tol)
        torch.testing.assert_close(pred(g["xi"].to(cuda), g["xj"].to(cuda)).float().cpu(), g["pred_out"], **tol)
        out3 = pred(g["x3i"].to(cuda), g["x3j"].to(cuda))
        assert out3.shape == g["pred_out3"].shape
        torch.testing.assert_close(out3.float().cpu(), g["pred_out3"], **tol)
        torch.testing.assert_close(pred_in(g["xi"].to(cuda), g["xj"].to(cuda)).float().cpu(), g["pred_in_out"],
                                   **(tol if mode == torch.float32 else dict(rtol=5e-2, atol=5e-2)))
        # fused scoring path == forward(h[u], h[v])
        h = sage(x, ei)
        u, v = ei[0, :64].contiguous(), ei[1, :64].contiguous()
        a = pred.score(h, u, v)
        b = pred(h[u].float(), h[v].float())
        torch.testing.assert_close(a.float(), b.float(), **tol)


def test_encoder_gradients_match_oracle(cuda, mode):
    """Full backward (SpMM transpose, weight-gradient GEMMs, dual input-gradient GEMM, gate) vs CPU autograd."""
    seed_all(0)
    n, f, hdim = 300, 40, 32
    ei = O.synthetic_undirected_graph(n, 1200, seed=2)
    x = torch.randn(n, f)
    for conv_o, conv_d in ((O.SAGEConv, L.SAGEConv), (O.SAGEConvUpdated, L.SAGEConv_updated)):
        mo = O.SAGE("c", f, hdim, hdim, 3, 0.0, conv_o)
        po = O.LinkPredictor("mlp", hdim, hdim, 1, 2, 0.0)
        md = L.SAGE("c", f, hdim, hdim, 3, 0.0, conv_d); md.load_state_dict(mo.state_dict()); md.to(cuda)
        pd = L.LinkPredictor("mlp", hdim, hdim, 1, 2, 0.0); pd.load_state_dict(po.state_dict()); pd.to(cuda)
        u, v = torch.randint(0, n, (500,)), torch.randint(0, n, (500,))
        lo = O.bce_loss(po(mo(x, ei)[u], mo(x, ei)[v]).squeeze(), torch.cat((torch.ones(200), torch.zeros(300))))
        lo.backward()
        hd = md(x.to(cuda), ei.to(cuda))
        ld = ops.bce_loss(pd.score(hd, u.to(cuda), v.to(cuda)).reshape(-1), 200)
        ld.backward()
        torch.testing.assert_close(ld.cpu(), lo.detach(), **TOL[mode])
        gtol = dict(rtol=1e-4, atol=1e-6) if mode == torch.float32 else dict(rtol=5e-2, atol=2e-3)
        for (k, a), (_, b) in zip(list(mo.named_parameters()) + list(po.named_parameters()),
                                  list(md.named_parameters()) + list(pd.named_parameters())):
            torch.testing.assert_close(b.grad.cpu(), a.grad, msg=lambda m, k=k: f"{k}: {m}", **gtol)


def test_wide_input_sageconv_updated_stacked_gemm(cuda, mode):
    """Coauthor-Physics-shaped first layer (C3): with >= 1024 input features the bf16 path computes lin_l and lin_r as
    ONE stacked GEMM + aggregate + epilogue pass.  One layer (no relu in between, so no mask flips): forward and all
    three parameter gradients vs the CPU oracle over two optimiser steps; then the 2-layer model forward.  Odd width
    (rows padded for TMA) and isolated nodes (the lin_l bias must not reach them)."""
    seed_all(1)
    n, f, hdim = 257, 1433, 32
    ei = O.synthetic_undirected_graph(n - 7, 900, seed=3)  # the last 7 nodes are isolated
    x = (torch.rand(n, f) < 0.02).float()
    co = O.SAGEConvUpdated(f, hdim)
    cd = L.SAGEConv_updated(f, hdim); cd.load_state_dict(co.state_dict()); cd.to(cuda)
    opt_o = torch.optim.Adam(co.parameters(), lr=0.01)
    opt_d = L.FusedAdam(cd.parameters(), lr=0.01)
    w = torch.randn(n, hdim)
    tol = TOL[mode]
    for step in range(2):
        opt_o.zero_grad(); opt_d.zero_grad()
        if mode == torch.bfloat16:
            # Adam's first steps move every weight by ~lr whatever the gradient size, so bf16 trajectories drift apart by
            # sign flips of tiny gradients: compare each step from the DEVICE's current weights (this also proves that
            # the stacked bf16 copies were refreshed by the optimiser step)
            co.load_state_dict({k: v.detach().cpu() for k, v in cd.state_dict().items()})
        ho = co(x, ei)
        hd = cd(x.to(cuda), ei.to(cuda))
        torch.testing.assert_close(hd.float().cpu(), ho.detach(), **tol)
        (ho * w).sum().backward()
        (hd.float() * w.to(cuda)).sum().backward()
        for (k, a), (_, b) in zip(co.named_parameters(), cd.named_parameters()):
            if mode == torch.float32:
                torch.testing.assert_close(b.grad.cpu(), a.grad, msg=lambda m, k=k: f"step {step} {k}: {m}", rtol=1e-4, atol=1e-5)
            else:  # bound the error against the size of the whole gradient (elements are sums of +/- terms that cancel)
                rel = float((b.grad.cpu() - a.grad).norm() / a.grad.norm())
                assert rel < 1e-2, (step, k, rel)
        opt_o.step(); opt_d.step()
    mo = O.SAGE("p", f, hdim, hdim, 2, 0.0, O.SAGEConvUpdated)
    md = L.SAGE("p", f, hdim, hdim, 2, 0.0, L.SAGEConv_updated); md.load_state_dict(mo.state_dict()); md.to(cuda).eval()
    with torch.no_grad():
        hs = md(x.to(cuda), ei.to(cuda))
        torch.testing.assert_close(hs.float().cpu(), mo.eval()(x, ei), **tol)
    if mode == torch.bfloat16:
        S = ops.stacked_weights(cd.lin_l.weight, cd.lin_r.weight)
        assert S is not None and S.shape == (2 * hdim, f)
        # the stacked halves are the live bf16 working copies of the two parameters (refreshed by the optimiser step)
        torch.testing.assert_close(S[:hdim].float(), cd.lin_l.weight.detach().bfloat16().float())
        torch.testing.assert_close(S[hdim:].float(), cd.lin_r.weight.detach().bfloat16().float())
        # stacked == unstacked path up to one bf16 rounding of lin_r x
        saved, ops.STACK_MIN_IN_FEATURES = ops.STACK_MIN_IN_FEATURES, 1 << 30
        try:
            with torch.no_grad():
                hu = md(x.to(cuda), ei.to(cuda))
        finally:
            ops.STACK_MIN_IN_FEATURES = saved
        torch.testing.assert_close(hs.float(), hu.float(), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("tag", ["teacher_fullbatch", "teacher_minibatch"])
def test_teacher_epochs_match_reference_golden(cuda, golden, mode, tag):
    g = golden[tag]
    x, split, H = g["x"], g["split"], g["H"]
    data = shims.Data(x=x, adj_t=split["train"]["edge"].t().contiguous()).to(cuda)
    model = L.SAGE("cora", x.size(1), H, H, 2, 0.0, L.SAGEConv); model.load_state_dict(g["sd0"]["gnn"]); model.to(cuda)
    predictor = L.LinkPredictor("mlp", H, H, 1, 2, 0.0); predictor.load_state_dict(g["sd0"]["predictor"]); predictor.to(cuda)
    opt = L.FusedAdam(list(model.parameters()) + list(predictor.parameters()), lr=g["lr"])
    seed_all(g["seed_train"])
    losses = [teacher.train(model, predictor, data, split, opt, g["batch_size"], "sage", "cora", "transductive") for _ in range(2)]
    rt = 1e-5 if mode == torch.float32 else 2e-2
    np.testing.assert_allclose(losses, g["losses"], rtol=rt)
    args = type("A", (), {"minibatch": False, "compute_auc": True})()
    results, h = teacher.test_transductive(model, predictor, data, split, L.Evaluator("ogbl-ddi"), g["batch_size"], "sage",
                                           "cora", args)
    if mode == torch.float32:  # (bf16: Adam moves every weight by ~lr per step whatever the gradient size, so
        # post-training embeddings are only comparable through the losses above)
        torch.testing.assert_close(h.float().cpu(), g["h"], rtol=1e-3, atol=1e-5)
        for K in (10, 20, 30, 50):  # reference-matching Hits@K (scores agree to ~1e-6, no near-ties in this fixture)
            assert results[f"Hits@{K}"] == pytest.approx(g["results"][f"Hits@{K}"], abs=1e-12)
        assert results["AUC"] == pytest.approx(g["results"]["AUC"], abs=5e-4)  # device pair counts (llp_auc_pairs) vs the reference's sklearn; one flipped pair = 8e-5
        for k, v in model.state_dict().items():
            torch.testing.assert_close(v.cpu(), g["sd1"]["gnn"][k], rtol=1e-3, atol=1e-5)


def test_student_epochs_match_reference_golden(cuda, golden, mode):
    g = golden["student"]
    a = type("A", (), dict(g["args"]))()
    x, split, H = g["x"], g["split"], g["H"]
    data = shims.Data(x=x, adj_t=split["train"]["edge"].t().contiguous()).to(cuda)
    model = L.MLP(2, x.size(1), H, H, 0.0); model.load_state_dict(g["sd0"]["mlp"]); model.to(cuda)
    pred = L.LinkPredictor("mlp", H, H, 1, 2, 0.0); pred.load_state_dict(g["sd0"]["predictor"]); pred.to(cuda)
    t_pred = L.LinkPredictor("mlp", H, H, 1, 2, 0.0); t_pred.load_state_dict(g["teacher_pred_sd"]); t_pred.to(cuda)
    for p in t_pred.parameters():
        p.requires_grad = False
    opt = L.FusedAdam(list(model.parameters()) + list(pred.parameters()), lr=g["lr"])
    shims.draw_rand_on_host(True)  # consume the CPU generator like the reference run that produced the fixture
    try:
        seed_all(g["seed_train"])
        losses = [student.train(model, pred, g["t_h"].to(cuda), t_pred, data, split, opt, a, cuda) for _ in range(2)]
    finally:
        shims.draw_rand_on_host(False)
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-5 if mode == torch.float32 else 2e-2)
    if mode == torch.float32:
        for k, v in model.state_dict().items():
            torch.testing.assert_close(v.cpu(), g["sd1"]["mlp"][k], rtol=1e-3, atol=1e-5)


def test_production_setting_matches_oracle(cuda, mode):
    """Production split (splits.do_production_edge_split) -> train() on the old-node graph ->
    test_production() over the five (positive, negative) pairs: loss, embeddings, Hits@K and AUC vs the CPU oracle."""
    from linkless_link_prediction_b200.data import synthetic_full_graph
    from linkless_link_prediction_b200.splits import do_production_edge_split
    full = synthetic_full_graph("cora", seed=0, scale=0.25)
    full.x = full.x[:, :64].contiguous()
    parts = do_production_edge_split([full], "cora", 0.3, 0.3, 0.3, 0.1)
    training_data, val_data, inference_data, _, bundle, negs = parts
    f, H = training_data.x.size(1), 32
    seed_all(0)
    mo = O.SAGE("cora", f, H, H, 2, 0.0); po = O.LinkPredictor("mlp", H, H, 1, 2, 0.0)
    md = L.SAGE("cora", f, H, H, 2, 0.0); md.load_state_dict(mo.state_dict()); md.to(cuda)
    pd = L.LinkPredictor("mlp", H, H, 1, 2, 0.0); pd.load_state_dict(po.state_dict()); pd.to(cuda)
    ro, ho = O.test_production(mo, po, val_data, inference_data, bundle, negs, 256)
    import copy
    dev_parts = [copy.copy(d).to(cuda) for d in (training_data, val_data, inference_data)]
    rd, hd = teacher.test_production(md, pd, dev_parts[1], dev_parts[2], bundle, negs, L.Evaluator(), 256, "sage", "cora")
    torch.testing.assert_close(hd.float().cpu(), ho, **TOL[mode])
    assert set(rd) == set(ro) and all(len(v) == 5 for v in rd.values())
    if mode == torch.float32:
        n_pos = [int(val_data.edge_label.sum()), bundle[3].size(1), bundle[0].size(1), bundle[1].size(1), bundle[2].size(1)]
        for K in (10, 20, 30, 50):  # at most one positive on the other side of a threshold (scores agree to ~1e-6)
            for a, b, n in zip(rd[f"Hits@{K}"], ro[f"Hits@{K}"], n_pos):
                assert abs(a - b) <= 1.0 / max(n, 1) + 1e-12
        assert rd["AUC"] == pytest.approx(ro["AUC"], abs=2e-3)
    # one training epoch on the training graph (production branch of train(): positives = its edge_index)
    opt_o = torch.optim.Adam(list(mo.parameters()) + list(po.parameters()), lr=0.01)
    opt_d = L.FusedAdam(list(md.parameters()) + list(pd.parameters()), lr=0.01)
    mo.train(); po.train()
    pos = training_data.edge_index
    g = torch.Generator().manual_seed(3)
    neg = torch.randint(0, training_data.x.size(0), pos.size(), generator=g)
    lo = O.teacher_step(mo, po, training_data.x, pos, pos, neg, opt_o)
    md.train(); pd.train()
    ld = teacher.train_step(md, pd, dev_parts[0], pos.to(cuda), neg.to(cuda), opt_d, "sage", "production").item()
    assert ld == pytest.approx(lo, rel=1e-5 if mode == torch.float32 else 2e-2)


def test_production_driver_runs_end_to_end(cuda, tmp_path, monkeypatch):
    """`python train_teacher_gnn.py --transductive=production` (scripts/supervised_production.sh): split generation,
    training epochs, five-way evaluation, result / checkpoint files in the reference's ../ layout."""
    work = tmp_path / "src"
    work.mkdir()
    monkeypatch.chdir(work)
    teacher.main(["--datasets=cora", "--encoder=sage", "--transductive=production", "--runs=1", "--epochs=2",
                  "--synthetic_scale=0.2", "--hidden_channels=32", "--batch_size=512", "--precision=fp32"])
    out = (tmp_path / "results" / "cora_supervised_production.txt").read_text()
    assert "All runs:" in out and "Final new_new" in out and "AUC" in out
    ck = torch.load(tmp_path / "saved-models" / "cora-sage_production.pkl", weights_only=False)
    assert set(ck) == {"gnn", "predictor"} and "convs.0.lin_l.weight" in ck["gnn"]
    feats = torch.load(tmp_path / "saved-features" / "cora-sage_production.pkl", weights_only=False)
    assert set(feats) == {"features"}
    ops.set_compute_dtype(torch.bfloat16)


def test_student_minibatch_equals_fullbatch_losses(cuda):
    """train_minibatch encodes only the touched rows; with dropout 0 its loss equals the full-batch step's."""
    ops.set_compute_dtype(torch.float32)
    try:
        n, f, H = 200, 24, 32
        ei = O.synthetic_undirected_graph(n, 700, seed=4)
        split = {"train": {"edge": ei.t().contiguous()}}
        data = shims.Data(x=torch.randn(n, f), adj_t=ei).to(cuda)
        t_h = torch.randn(n, H).to(cuda)
        args = type("A", (), dict(transductive="transductive", node_batch_size=50, link_batch_size=400, LLP_R=1.0,
                                  LLP_D=1.0, True_label=1.0, KD_RM=0.0, KD_LM=0.0, margin=0.1, rw_step=2, ps_method="nb",
                                  ns_rate=2, hops=2, datasets="cora"))()
        outs = []
        for fn in (student.train, student.train_minibatch):
            seed_all(3)
            model = L.MLP(2, f, H, H, 0.0).to(cuda)
            pred = L.LinkPredictor("mlp", H, H, 1, 2, 0.0).to(cuda)
            t_pred = L.LinkPredictor("mlp", H, H, 1, 2, 0.0).to(cuda)
            opt = L.FusedAdam(list(model.parameters()) + list(pred.parameters()), lr=0.01)
            shims.draw_rand_on_host(True)
            seed_all(4)
            outs.append(fn(model, pred, t_h, t_pred, data, split, opt, args, cuda))
            shims.draw_rand_on_host(False)
        assert outs[0] == pytest.approx(outs[1], rel=1e-5)
    finally:
        ops.set_compute_dtype(torch.bfloat16)
        shims.draw_rand_on_host(False)


def test_minibatch_every_node_once_equals_gathered_rows(cuda):
    """student_minibatch_step encodes every node ONCE when the step touches >= N rows and the encoder is deterministic
    (main.encode_every_node_once) instead of the gathered rows with their duplicates (main.py:93-101): same epoch loss and
    the same parameters after the epoch as the gathered-row path; with dropout the rule must not apply."""
    ops.set_compute_dtype(torch.float32)
    try:
        n, f, H = 300, 24, 32
        ei = O.synthetic_undirected_graph(n, 900, seed=5)
        split = {"train": {"edge": ei.t().contiguous()}}
        data = shims.Data(x=torch.randn(n, f), adj_t=ei).to(cuda)
        t_h = torch.randn(n, H).to(cuda)
        args = type("A", (), dict(transductive="transductive", node_batch_size=40, link_batch_size=300, LLP_R=1.0,
                                  LLP_D=1.0, True_label=1.0, KD_RM=0.0, KD_LM=0.0, margin=0.1, rw_step=2, ps_method="nb",
                                  ns_rate=2, hops=2, datasets="cora"))()   # 8 node batches >= the link batches of the epoch
        res = []
        for once in (True, False):
            student.ENCODE_EVERY_NODE_ONCE = once
            seed_all(3)
            model = L.MLP(3, f, H, H, 0.0).to(cuda)
            pred = L.LinkPredictor("mlp", H, H, 1, 3, 0.0).to(cuda)
            t_pred = L.LinkPredictor("mlp", H, H, 1, 2, 0.0).to(cuda)
            opt = L.FusedAdam(list(model.parameters()) + list(pred.parameters()), lr=0.01)
            assert student.encode_every_node_once(model, n, 10 * n) == once
            shims.draw_rand_on_host(True)
            seed_all(4)
            loss = student.train_minibatch(model, pred, t_h, t_pred, data, split, opt, args, cuda)
            shims.draw_rand_on_host(False)
            res.append((loss, [q.detach().clone() for q in list(model.parameters()) + list(pred.parameters())]))
        assert res[0][0] == pytest.approx(res[1][0], rel=1e-5)
        for a, b in zip(res[0][1], res[1][1]):
            torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-5)
        student.ENCODE_EVERY_NODE_ONCE = True
        drop = L.MLP(3, f, H, H, 0.5).to(cuda).train()
        assert not student.encode_every_node_once(drop, n, 10 * n)      # active dropout: one mask per occurrence upstream
        assert student.encode_every_node_once(drop.eval(), n, 10 * n)
        assert not student.encode_every_node_once(L.MLP(3, f, H, H, 0.0), n, n - 1)   # fewer rows than nodes: gather
    finally:
        student.ENCODE_EVERY_NODE_ONCE = True
        ops.set_compute_dtype(torch.bfloat16)
        shims.draw_rand_on_host(False)


def test_input_aggregation_cache_is_bit_identical_and_invalidated(cuda):
    """ops.Graph.spmm_input keeps the aggregate of the constant input features on the graph (layer 1 of SAGE over PyG
    SAGEConv): training steps with and without it give bit-identical embeddings, losses and parameters; an in-place change
    of the features (version counter) or another feature tensor recomputes it; hidden layers never use it."""
    data, split = L.data.synthetic_dataset("cora", seed=0, scale=0.5)
    data = data.to(cuda)
    pos = split["train"]["edge"].to(cuda)[:1500].t().contiguous()
    neg = torch.randint(0, data.x.size(0), pos.size(), device=cuda)
    outs = []
    for cached in (True, False):
        ops.CACHE_INPUT_AGGREGATION = cached
        seed_all(1)
        model = L.SAGE("cora", data.x.size(1), 64, 64, 3, 0.5).to(cuda)
        pred = L.LinkPredictor("mlp", 64, 64, 1, 2, 0.5).to(cuda)
        opt = L.FusedAdam(list(model.parameters()) + list(pred.parameters()), lr=0.01)
        model.train(); pred.train()
        ops.seed_dropout(5)
        graph = ops.graph_of(data.adj_t, data.x.size(0))
        graph.__dict__.pop("_input_agg", None)
        losses = [float(teacher.train_step(model, pred, data, pos, neg, opt)) for _ in range(3)]
        assert (getattr(graph, "_input_agg", None) is not None) == cached
        model.eval()
        with torch.no_grad():
            h = model(data.x, data.adj_t)
        outs.append((losses, h.clone(), [q.detach().clone() for q in model.parameters()]))
    ops.CACHE_INPUT_AGGREGATION = True
    assert outs[0][0] == outs[1][0]
    assert torch.equal(outs[0][1], outs[1][1])
    assert all(torch.equal(a, b) for a, b in zip(outs[0][2], outs[1][2]))
    # invalidation
    g = ops.graph_of(data.adj_t, data.x.size(0))
    xc = ops.to_compute(data.x.clone())
    a1 = g.spmm_input(xc)
    assert g.spmm_input(xc) is a1
    xc.mul_(2.0)                                   # in-place change: version counter moves
    a2 = g.spmm_input(xc)
    assert a2 is not a1 and torch.equal(a2, g.spmm(xc))
    other = ops.to_compute(torch.randn_like(data.x))
    assert torch.equal(g.spmm_input(other), g.spmm(other))


def test_training_with_dropout_learns(cuda):
    """Dropout path (fused Philox epilogue + gate backward): the loss must go down on a learnable toy problem."""
    seed_all(0)
    data, split = L.data.synthetic_dataset("cora", seed=0, scale=0.25)
    data = data.to(cuda)
    model = L.SAGE("cora", data.x.size(1), 64, 64, 2, 0.5).to(cuda)
    pred = L.LinkPredictor("mlp", 64, 64, 1, 2, 0.5).to(cuda)
    opt = L.FusedAdam(list(model.parameters()) + list(pred.parameters()), lr=0.01)
    losses = [teacher.train(model, pred, data, split, opt, 65536, "sage", "cora", "transductive") for _ in range(30)]
    assert losses[-1] < 0.8 * losses[0]
    assert all(np.isfinite(losses))


def test_fused_gate_equals_explicit_gate(cuda):
    """The relu/dropout backward mask applied in the consumer's GEMM epilogue (in_gate / defer_gate) must give the
    same gradients as the stand-alone gate kernel, with dropout active (same Philox seeds => same masks)."""
    ops.set_compute_dtype(torch.float32)
    try:
        seed_all(1)
        n, f, H = 400, 32, 64
        ei = O.synthetic_undirected_graph(n, 1500, seed=3).to(cuda)
        x = torch.randn(n, f).to(cuda)
        c1, c2 = L.SAGEConv(f, H).to(cuda), L.SAGEConv(H, H).to(cuda)
        lin = torch.nn.Linear(H, 1).to(cuda)
        grads = []
        for fused in (True, False):
            for m in (c1, c2, lin):
                m.zero_grad()
            c1.train(); c2.train()
            ops.seed_dropout(2)  # identical dropout streams (key, step, site ids) in both variants
            h1 = c1(x, ei, _relu=True, _dropout=0.4, _in_gate=0.0, _defer_gate=fused)
            h2 = c2(h1, ei, _relu=True, _dropout=0.4, _in_gate=(1.0 / 0.6) if fused else 0.0, _defer_gate=fused)
            p = ops.ScoreHeadFn.apply(h2, lin.weight, lin.bias, (1.0 / 0.6) if fused else 0.0)
            ops.bce_loss(p, n // 2).backward()
            grads.append([q.grad.clone() for m in (c1, c2, lin) for q in m.parameters()])
        for a, b in zip(*grads):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)
    finally:
        ops.set_compute_dtype(torch.bfloat16)


def _teacher_pair(cuda, dropout, seed=0, n=600, f=48, H=64):
    seed_all(seed)
    ei = O.synthetic_undirected_graph(n, 2500, seed=5)
    data = shims.Data(x=torch.randn(n, f), adj_t=ei).to(cuda)
    model = L.SAGE("c", f, H, H, 2, dropout).to(cuda)
    pred = L.LinkPredictor("mlp", H, H, 1, 2, dropout).to(cuda)
    model.train(); pred.train()
    return data, ei, model, pred


def test_captured_step_matches_eager_steps(cuda, mode):
    """CUDA-graph replays of the training step are the same optimisation steps as eager calls (dropout 0): same
    losses, same parameters, including Adam's bias correction from the device-side step counter."""
    batches = None
    outs = []
    for captured in (False, True):
        data, ei, model, pred = _teacher_pair(cuda, 0.0)
        opt = L.FusedAdam(list(model.parameters()) + list(pred.parameters()), lr=0.01)
        if batches is None:
            g = torch.Generator().manual_seed(7)
            batches = [(ei[:, torch.randint(0, ei.size(1), (512,), generator=g)].to(cuda),
                        torch.randint(0, 600, (2, 512), generator=g).to(cuda)) for _ in range(7)]
        step = teacher.CapturedTrainStep(model, pred, data, opt, eager_steps=2) if captured else \
            (lambda e, n_: teacher.train_step(model, pred, data, e, n_, opt))
        losses = [float(step(e, n_).item()) for e, n_ in batches]
        if captured:
            assert step.graph is not None and step.replays == 5 and step.launches_per_replay > 10
        outs.append((losses, [p.detach().clone() for p in list(model.parameters()) + list(pred.parameters())]))
    rt = 1e-5 if mode == torch.float32 else 2e-2
    np.testing.assert_allclose(outs[1][0], outs[0][0], rtol=rt)
    for a, b in zip(outs[0][1], outs[1][1]):
        torch.testing.assert_close(b, a, rtol=1e-3, atol=1e-4 if mode == torch.float32 else 3e-3)


def test_captured_step_draws_new_dropout_masks(cuda):
    """The dropout stream is keyed on a device-side step counter, so every replay of the captured graph sees new
    masks: with lr = 0 (parameters frozen) and one fixed batch, the per-replay losses must differ from each other,
    and re-seeding must reproduce the sequence."""
    data, ei, model, pred = _teacher_pair(cuda, 0.5)
    opt = L.FusedAdam(list(model.parameters()) + list(pred.parameters()), lr=0.0)
    e, n_ = ei[:, :512].contiguous().to(cuda), torch.randint(0, 600, (2, 512)).to(cuda)
    step = teacher.CapturedTrainStep(model, pred, data, opt, eager_steps=1)
    shims.seed_everything(11)
    a = [float(step(e, n_).item()) for _ in range(6)]
    assert step.replays == 5
    assert len({round(v, 7) for v in a[1:]}) == 5, a
    ops.seed_dropout(11)
    b = [float(step(e, n_).item()) for _ in range(5)]
    # replays 0..4 after re-seeding use device steps 1..5 again -> a[0] was the eager step 1 with other site ids
    assert len({round(v, 7) for v in b}) == 5
    ops.seed_dropout(11)
    c = [float(step(e, n_).item()) for _ in range(5)]
    np.testing.assert_allclose(b, c, rtol=1e-6)


@pytest.mark.parametrize("p", [0.0, 0.5, 0.3])
@pytest.mark.parametrize("n,f,H,m", [(500, 64, 64, 1000), (3000, 256, 256, 20000), (700, 128, 200, 777)])
def test_fused_edge_scorer_equals_unfused(cuda, p, n, f, H, m):
    """LinkPredictor.score through the single fused kernel (gather-Hadamard as the A-operand producer of the first
    predictor GEMM + bias/relu/dropout + score head) == the separate kernels: same dropout masks, same bf16 roundings of
    z and y, scores and every gradient agree; the no-grad (evaluation) path writes only the scores."""
    from linkless_link_prediction_b200 import models
    seed_all(3)
    h0 = (torch.randn(n, f) * 0.5).to(cuda)
    u, v = torch.randint(0, n, (m,)).to(cuda), torch.randint(0, n, (m,)).to(cuda)
    pred = L.LinkPredictor("mlp", f, H, 1, 2, p).to(cuda)
    pred.train()
    dp = torch.randn(m).to(cuda)
    outs = []
    for fused in (True, False):
        models.FUSE_EDGE_MLP = fused
        try:
            for q in pred.parameters():
                q.grad = None
            ops.seed_dropout(5)
            ops.advance_rng(cuda)
            h = ops.cast2d(h0, torch.bfloat16).requires_grad_(True)
            prob = pred.score(h, u, v).reshape(-1)
            prob.backward(dp)
            with torch.no_grad():
                pred.eval()
                ev = pred.score(h.detach(), u, v).reshape(-1)
                pred.train()
            outs.append((prob.detach(), h.grad.float(), [q.grad.clone() for q in pred.parameters()], ev))
        finally:
            models.FUSE_EDGE_MLP = True
    (pa, ga, qa, ea), (pb, gb, qb, eb) = outs
    torch.testing.assert_close(pa, pb, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ea, eb, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ga, gb, rtol=2e-2, atol=2e-3)
    for a, b in zip(qa, qb):
        torch.testing.assert_close(a, b, rtol=1e-3, atol=1e-3)
    # and against the fp32 oracle formula at p = 0
    if p == 0.0:
        w1, b1, w2, b2 = [q.detach().float().cpu() for q in pred.parameters()]
        hf = ops.cast2d(h0, torch.bfloat16).float().cpu()
        z = hf[u.cpu()] * hf[v.cpu()]
        ref = torch.sigmoid(torch.relu(z @ w1.t() + b1) @ w2.t() + b2).reshape(-1)
        torch.testing.assert_close(ea.cpu(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("norm_type", ["batch", "layer"])
def test_norm_layers_and_l2_normalize_match_torch_formula(cuda, mode, norm_type):
    """The surface options the reference drivers never enable but its classes accept: ``norm_type`` of MLP / SAGE
    (models.py:27-37,48-53,90-100,113-116) and ``SAGEConv_updated(normalize=True)`` (sageconv_updated.py:78-79): forward
    and parameter gradients against the same formulas written with plain torch + the oracle's aggregation on the CPU."""
    import torch.nn.functional as F
    seed_all(0)
    n, f, H = 300, 24, 32
    ei = O.synthetic_undirected_graph(n, 1200, seed=2)
    x = torch.randn(n, f)
    tol = TOL[mode] if mode == torch.float32 else dict(rtol=5e-2, atol=5e-2)

    def check(dev_model, ref_forward, dev_inputs):
        dev_model.train()
        ref_params = [p.detach().cpu().clone().requires_grad_(True) for p in dev_model.parameters()]
        out_ref = ref_forward(ref_params)
        w = torch.randn(out_ref.shape)
        (out_ref * w).sum().backward()
        out = dev_model(*dev_inputs)
        (out.float() * w.to(cuda)).sum().backward()
        torch.testing.assert_close(out.float().cpu(), out_ref.detach(), **tol)
        # the bias in front of a batch norm has an analytically ZERO gradient (the norm removes the mean): errors are
        # bounded against the larger of the tensor's own norm and 1e-3 (bf16: 2e-2) of the largest gradient norm of the model
        scale = max(float(q.grad.norm()) for q in ref_params)
        for p, q in zip(dev_model.parameters(), ref_params):
            err = float((p.grad.cpu() - q.grad).norm()) / max(float(q.grad.norm()), (1e-3 if mode == torch.float32 else 2e-2) * scale)
            assert err < (1e-4 if mode == torch.float32 else 1.5e-1), err

    norm = (lambda h, wgt, b: F.batch_norm(h, None, None, wgt, b, True)) if norm_type == "batch" else \
        (lambda h, wgt, b: F.layer_norm(h, (H,), wgt, b))
    # MLP 3 layers, dropout 0: linear -> norm -> relu, twice, then linear
    mlp = L.MLP(3, f, H, 16, 0.0, norm_type).to(cuda)
    names = [k for k, _ in mlp.named_parameters()]

    def mlp_ref(ps):
        P = dict(zip(names, ps))
        h = x
        for l in range(3):
            h = F.linear(h, P[f"layers.{l}.weight"], P[f"layers.{l}.bias"])
            if l != 2:
                h = torch.relu(norm(h, P[f"norms.{l}.weight"], P[f"norms.{l}.bias"]))
        return h

    check(mlp, mlp_ref, (x.to(cuda),))
    # SAGE 3 layers with norms, the last conv with L2 normalisation
    sage = L.SAGE("t", f, H, H, 3, 0.0, L.SAGEConv_updated, norm_type).to(cuda)
    sage.convs[2].normalize = True
    snames = [k for k, _ in sage.named_parameters()]

    def sage_ref(ps):
        P = dict(zip(snames, ps))
        h = x
        for l in range(3):
            t = F.linear(h, P[f"convs.{l}.lin_l.weight"], P[f"convs.{l}.lin_l.bias"])
            h = O.mean_aggregate(t, ei, n) + F.linear(h, P[f"convs.{l}.lin_r.weight"])
            if l == 2:
                h = F.normalize(h, p=2.0, dim=-1)
            else:
                h = torch.relu(norm(h, P[f"norms.{l}.weight"], P[f"norms.{l}.bias"]))
        return h

    check(sage, sage_ref, (x.to(cuda), ei.to(cuda)))


@pytest.mark.parametrize("setting", ["transductive", "production"])
def test_teacher_then_student_drivers_run_end_to_end(cuda, tmp_path, monkeypatch, setting):
    """scripts/supervised_*.sh followed by scripts/LLP_*.sh: the teacher driver writes ../saved-models and
    ../saved-features, the student driver (main.py) loads them through the reference's file names, trains with
    LLP_D + LLP_R + True_label on the SAME split (one shared loader) and writes its result file — both settings."""
    work = tmp_path / "src"
    work.mkdir()
    monkeypatch.chdir(work)
    common = ["--datasets=cora", "--encoder=sage", f"--transductive={setting}", "--runs=1", "--epochs=2",
              "--synthetic_scale=0.2", "--precision=fp32"]
    teacher.main(common + ["--hidden_channels=256", "--batch_size=512"])   # the student hard-codes a 256-wide teacher predictor
    assert (tmp_path / "saved-models" / f"cora-sage_{setting}.pkl").exists()
    student.main(common + ["--hidden_channels=256", "--link_batch_size=512", "--LLP_D=1", "--LLP_R=1", "--True_label=1",
                           "--dropout=0.0", "--rw_step=2", "--hops=2", "--ns_rate=1"])
    out = (tmp_path / "results" / f"cora_KD_{setting}.txt").read_text()
    assert "LLP (Relational Distillation)" in out and "All runs:" in out and "AUC" in out
    if setting == "production":
        assert "Final new_new" in out
    ops.set_compute_dtype(torch.bfloat16)
