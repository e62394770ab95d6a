"""Host-side logic, the CLI surface and the C-ABI export table — runs without a GPU."""
import ctypes
import os
import re

import pytest
import torch

import linkless_link_prediction_b200 as L
from linkless_link_prediction_b200 import _native as N
from linkless_link_prediction_b200 import data as D
from linkless_link_prediction_b200 import loader, main as student, train_teacher_gnn as teacher
from oracle import llp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = N.load()
    header = open(os.path.join(ROOT, "include", "llp_b200.h")).read()
    declared = set(re.findall(r"\b(llp_[a-z0-9_]+)\s*\(", header))
    declared.discard("llp_gemm_nt_args")
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/llp_b200.h but not exported"
        assert name in N.PROTOTYPES, f"{name} has no ctypes prototype"
    assert lib.llp_version() == 100
    assert b"no fallback" in lib.llp_error_string(-4)
    # struct layout must match the C definition (8-byte fields after four ints)
    assert ctypes.sizeof(N.GemmNtArgs) == 4 * 4 + 8 * 4 + 8 * 8 + 8 + 16 + 16 + 8 + 16 + 8 + 16


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.ops.gemm_nt(torch.zeros(4, 8), torch.zeros(4, 8))
    model = L.SAGE("cora", 8, 16, 16, 2, 0.0)
    with pytest.raises(RuntimeError):
        model(torch.zeros(5, 8), torch.zeros(2, 3, dtype=torch.long))
    # the rows added later in the round fail just as loudly: device AUC, the stand-alone epilogue, the partitioned graph
    with pytest.raises(RuntimeError):
        L.ops.auc_pairs(torch.rand(4), torch.rand(5))
    with pytest.raises(RuntimeError):
        L.shims.roc_auc_score_device(torch.rand(4), torch.rand(5))
    with pytest.raises(RuntimeError):
        L.ops.add_act(torch.zeros(4, 8), relu=True)
    with pytest.raises(RuntimeError):
        L.ops.PartitionedGraph(torch.zeros(2, 3, dtype=torch.long), 5, 0, 2)


@pytest.mark.parametrize("seed,n,b", [(0, 100, 32), (5, 8976, 65536), (7, 1000, 100), (9, 7, 3)])
def test_shuffled_batches_equal_dataloader(seed, n, b):
    from torch.utils.data import DataLoader
    torch.manual_seed(seed)
    ref = [p for p in DataLoader(range(n), b, shuffle=True)]
    after_ref = torch.rand(1)
    torch.manual_seed(seed)
    got = list(loader.shuffled_batches(n, b))
    after_got = torch.rand(1)
    assert len(ref) == len(got) and all(torch.equal(x, y) for x, y in zip(ref, got))
    assert torch.equal(after_ref, after_got)  # same number of global-RNG draws


def test_state_dict_keys_match_reference_layout(golden):
    sage = L.SAGE("cora", 24, 32, 16, 3, 0.5, L.SAGEConv)
    assert list(sage.state_dict().keys()) == list(golden["models"]["sage_sd"].keys())
    sage.load_state_dict(golden["models"]["sage_sd"], strict=True)
    sage_u = L.SAGE("coauthor-physics", 24, 32, 16, 2, 0.5, L.SAGEConv_updated)
    sage_u.load_state_dict(golden["models"]["sage_u_sd"], strict=True)
    L.MLP(3, 24, 32, 16, 0.5).load_state_dict(golden["models"]["mlp_sd"], strict=True)
    L.LinkPredictor("mlp", 16, 32, 1, 3, 0.5).load_state_dict(golden["models"]["pred_sd"], strict=True)


def test_parameter_init_stream_matches_oracle():
    torch.manual_seed(3)
    a = L.SAGE("cora", 12, 8, 8, 2, 0.5)
    pa = L.LinkPredictor("mlp", 8, 8, 1, 2, 0.5)
    torch.manual_seed(3)
    b = O.SAGE("cora", 12, 8, 8, 2, 0.5)
    pb = O.LinkPredictor("mlp", 8, 8, 1, 2, 0.5)
    for (k1, v1), (k2, v2) in zip(list(a.state_dict().items()) + list(pa.state_dict().items()),
                                  list(b.state_dict().items()) + list(pb.state_dict().items())):
        assert k1 == k2 and torch.equal(v1, v2)


def test_cli_flags_match_reference():
    t = teacher.build_parser().parse_args([])
    assert (t.device, t.num_layers, t.hidden_channels, t.dropout, t.batch_size, t.lr, t.epochs, t.runs, t.datasets,
            t.predictor, t.patience, t.transductive, t.encoder) == (0, 2, 256, 0.5, 65536, 0.005, 20000, 5, "cora",
                                                                    "mlp", 100, "transductive", "sage")
    s = student.build_parser().parse_args([])
    assert (s.True_label, s.KD_RM, s.KD_LM, s.LLP_D, s.LLP_R, s.margin, s.rw_step, s.ns_rate, s.hops, s.ps_method,
            s.link_batch_size, s.node_batch_size, s.runs, s.datasets) == (0.1, 0, 0, 1, 1, 0.1, 3, 1, 2, "nb", 65536,
                                                                          65536, 10, "collab")
    s = student.build_parser().parse_args("--datasets=cora --LLP_D=0.001 --LLP_R=1 --True_label=0.1 --minibatch".split())
    assert s.minibatch and s.LLP_D == 0.001


def test_synthetic_shapes():
    data, split = D.synthetic_dataset("cora", seed=0)
    assert data.x.shape == (2708, 1433)
    e = split["train"]["edge"]
    assert data.adj_t.shape[0] == 2 and data.adj_t.shape[1] == e.shape[0]
    assert e.shape[0] % 2 == 0 and abs(e.shape[0] - 8976) <= 4          # both directions of 85% of 5278 pairs
    assert split["valid"]["edge"].shape[0] == 263 and split["test"]["edge"].shape[0] == 527
    key = e[:, 0] * 2708 + e[:, 1]
    assert torch.equal(key, torch.sort(key).values)                      # (row, col)-sorted like to_undirected


def test_shard_partition_covers_everything():
    for n in (0, 1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            parts = [teacher._shard(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))


def test_logger_best_by_validation():
    lg = L.logger.Logger(1)
    for r in [(0.1, 0.5), (0.3, 0.2), (0.2, 0.9)]:
        lg.add_result(0, r)
    r, arg = lg.best(0)
    assert arg == 1 and r[arg, 1].item() == pytest.approx(20.0)
    with pytest.raises(AssertionError):
        L.logger.ProductionLogger(1).add_result(0, (1, 2))


def test_interleaved_loaders_consume_rng_like_reference():
    # main.py:167-171: the node loader iterator is created first, the link loader drives the loop
    from torch.utils.data import DataLoader
    torch.manual_seed(1)
    nl = iter(DataLoader(range(50), 7, shuffle=True))
    ref = [(lp, next(nl)) for lp in DataLoader(range(90), 40, shuffle=True)]
    torch.manual_seed(1)
    nl2 = loader.shuffled_batches(50, 7)
    got = [(lp, next(nl2)) for lp in loader.shuffled_batches(90, 40)]
    assert len(ref) == len(got)
    for (a, b), (c, d) in zip(ref, got):
        assert torch.equal(a, c) and torch.equal(b, d)


@pytest.mark.parametrize("world", [1, 3, 8])
def test_partition_messages_owns_every_message_once(world):
    """Node partition of the encoder (ops.partition_messages): blocks tile the padded node range, every
    message is owned exactly once per direction in its original order, the local aggregation of all ranks stacked
    equals the unpartitioned aggregation bit for bit."""
    from linkless_link_prediction_b200.ops import partition_messages
    from oracle import llp_oracle as O
    n = 103
    ei = O.synthetic_undirected_graph(n - 5, 700, seed=4)
    x = torch.randn(n, 6, generator=torch.Generator().manual_seed(1))
    parts = [partition_messages(ei, n, r, world) for r in range(world)]
    n_loc = parts[0][0]
    assert n_loc == -(-n // world) and [p[1] for p in parts] == [r * n_loc for r in range(world)]
    x_pad = torch.zeros(n_loc * world, 6); x_pad[:n] = x
    rp, col, _ = O.csr_build(ei, n, "dst")
    ref = O.spmm_csr(rp, col, x, mean=True)
    rows = []
    seen_in = seen_out = 0
    for n_loc_r, lo, hi, (f_src, f_dst), (t_src, t_dst), inv_deg in parts:
        assert f_dst.numel() == 0 or (int(f_dst.min()) >= 0 and int(f_dst.max()) < n_loc)
        assert t_src.numel() == 0 or int(t_src.max()) < n_loc
        own = (ei[1] >= lo) & (ei[1] < hi)
        assert torch.equal(f_src, ei[0][own]) and torch.equal(f_dst + lo, ei[1][own])      # original order kept
        seen_in += f_src.numel(); seen_out += t_src.numel()
        rp_l, col_l, _ = O.csr_build(torch.stack([f_src, f_dst]), n_loc, "dst")
        rows.append(O.spmm_csr(rp_l, col_l, x_pad, mean=True))
        deg = torch.zeros(n_loc * world).index_add_(0, ei[1], torch.ones(ei.size(1)))
        assert torch.equal(inv_deg, 1.0 / deg.clamp(min=1))
    assert seen_in == ei.size(1) and seen_out == ei.size(1)
    assert torch.equal(torch.cat(rows)[:n], ref)


def test_py_random_sample_is_bit_exact_with_cpython():
    """`llp_py_random_sample` (C++ restatement of CPython's `random.sample(range(n), k)`: MT19937, getrandbits,
    _randbelow rejection, pool vs set selection) returns the same indices AND leaves Python's global generator in the
    same state — the candidate stream of PyG's negative_sampling (train_teacher_gnn.py:50-51) stays bit-exact."""
    import random

    from linkless_link_prediction_b200 import shims
    cases = [(7_330_556, 9_873), (10, 10), (25, 6), (30, 5), (100, 3), (1_189_732_556, 72_089), (1 << 33, 1000),
             ((1 << 40) + 12345, 257), (4 ** 7 + 21, 4 ** 6), (4 ** 7 + 22, 4 ** 6), (2, 1), (65, 64), (1000, 999)]
    for seed, (n, k) in enumerate(cases):
        random.seed(seed)
        ref = random.sample(range(n), k)
        after_ref = random.random()
        random.seed(seed)
        got = shims.py_random_sample(n, k).tolist()
        after_got = random.random()
        assert got == ref, (n, k)
        assert after_got == after_ref, (n, k)
    # consecutive draws continue one stream
    random.seed(99)
    a = [random.sample(range(5000), 40) for _ in range(30)]
    random.seed(99)
    b = [shims.py_random_sample(5000, 40).tolist() for _ in range(30)]
    assert a == b


def test_speculative_candidate_draw_keeps_cpython_stream():
    """shims._candidates_pinned starts the NEXT step's random.sample on a worker thread; the result is only used (and
    Python's state only advanced) when the next request is identical and nobody touched `random` in between."""
    import random

    import torch

    from linkless_link_prediction_b200 import shims
    for seed in (1, 2):
        random.seed(seed)
        got = []
        for i in range(6):
            got.append(shims._candidates_pinned(7330556, 2000 if i != 4 else 2500, pin=False).clone())
            if i == 2:
                random.random()          # another user of the generator: the speculation must be dropped
        s_got = random.getstate()
        random.seed(seed)
        ref = []
        for i in range(6):
            ref.append(torch.tensor(random.sample(range(7330556), 2000 if i != 4 else 2500)))
            if i == 2:
                random.random()
        assert all(torch.equal(a, b) for a, b in zip(got, ref))
        assert s_got == random.getstate()
    random.seed(9); shims._candidates_pinned(10 ** 6, 300, pin=False)
    random.seed(9)                       # re-seeding between two calls: the pending draw belongs to the old stream
    a = shims._candidates_pinned(10 ** 6, 300, pin=False)
    random.seed(9)
    assert torch.equal(a, torch.tensor(random.sample(range(10 ** 6), 300)))


def test_negative_sampling_dense_equals_mask_formulation():
    """The sorted-id membership test that replaced PyG's N*N - N boolean mask keeps exactly the candidates the mask keeps
    (same order), including the second round after an incomplete first one."""
    import random

    import torch

    from linkless_link_prediction_b200 import shims
    from oracle import llp_oracle as O
    for n, pairs, want, seed in ((60, 200, 150, 3), (40, 700, 600, 4), (300, 2000, 1000, 5)):
        ei = O.synthetic_undirected_graph(n, pairs, seed=seed)
        random.seed(seed)
        ref = O.negative_sampling_dense(ei, n, want)
        random.seed(seed)
        got = shims.negative_sampling(ei, num_nodes=n, num_neg_samples=want, method="dense")
        assert torch.equal(got, ref), (n, pairs, want)
