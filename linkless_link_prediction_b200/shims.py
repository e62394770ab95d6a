"""Stand-ins for the third-party entry points the reference drivers import, with the same names,
arguments and error behaviour, backed by the B200 kernels (torch_geometric / torch_cluster / ogb are
not installed here and are not needed):

* ``negative_sampling``  — torch_geometric.utils.negative_sampling (train_teacher_gnn.py:50-51, main.py:81,206)
* ``random_walk``        — torch_cluster.random_walk (main.py:37,43,45)
* ``Evaluator``          — ogb.linkproppred.Evaluator (train_teacher_gnn.py:394,120-145)
* ``seed_everything``    — torch_geometric.seed.seed_everything (train_teacher_gnn.py:422, main.py:396)
* ``Data``               — the attribute bag the step functions read (.x .adj_t .edge_index .edge_label ...)
"""
from __future__ import annotations

import random
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import ops


def seed_everything(seed: int) -> None:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    ops.seed_dropout(seed)


class Data:
    """Minimal PyG ``Data`` look-alike: attribute container with ``.to(device)`` moving every tensor."""

    def __init__(self, x=None, edge_index=None, **kwargs):
        if x is not None:
            self.x = x
        if edge_index is not None:
            self.edge_index = edge_index
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    @property
    def num_nodes(self):
        return self.x.size(0)

    @property
    def num_features(self):
        return self.x.size(1)


def py_random_sample(population: int, k: int, pin: bool = False) -> torch.Tensor:
    """``torch.tensor(random.sample(range(population), k))`` — same values, same advance of Python's global ``random``
    state — through the C++ restatement of CPython's algorithm in ``libllp_b200.so`` (``llp_py_random_sample``; host code,
    works without a GPU).  The Python loop costs ~0.4 us per draw: 4-30 ms per training step at 10^4-10^5 candidates.
    ``pin``: return page-locked memory (for an asynchronous copy to the device)."""
    import array

    from . import _native as N
    version, internal, gauss = random.getstate()
    state = array.array("I", internal)   # 624 MT19937 words + the position; array <-> tuple is ~10x cheaper than via numpy
    out = torch.empty(k, dtype=torch.int64, pin_memory=bool(pin))
    N.check(N.load().llp_py_random_sample(state.buffer_info()[0], int(population), int(k), out.data_ptr()),
            "llp_py_random_sample")
    random.setstate((version, tuple(state), gauss))
    return out


def _sample_ids(population: int, k: int, device) -> torch.Tensor:
    """PyG 2.2.0 ``utils.negative_sampling.sample``: CPython ``random.sample`` on the host."""
    if population <= k:
        return torch.arange(population, device=device)
    return py_random_sample(population, k).to(device)


# ---- speculative candidate draw -------------------------------------------------------------------------------------
# The training loops call dense negative_sampling once per step with the same (population, sample size); the host's
# MT19937 draw (0.16 ms for Cora's 9.9k candidates, 1.1 ms for Physics' 72k) is the longest host-side piece of a step.
# After serving a draw, the NEXT one is started on a worker thread from the state the generator now has (the C++ call
# runs without the GIL).  It is purely speculative: Python's `random` state is NOT advanced until the next call consumes
# the result, and that call only does so if the request is the same and `random.getstate()` is still exactly the state
# the draw started from (nobody seeded or used `random` in between); otherwise the result is dropped and the draw
# happens in line.  Same values, same state after every call, as `random.sample(range(population), k)`.
PREFETCH_CANDIDATES = True
_PREFETCH = None      # (population, k, state tuple the draw started from, future -> (pinned tensor, state tuple after))
_PREFETCH_POOL = None


def _draw(state_words, population: int, k: int, out: torch.Tensor):
    import array

    from . import _native as N
    st = array.array("I", state_words)
    N.check(N.load().llp_py_random_sample(st.buffer_info()[0], int(population), int(k), out.data_ptr()), "llp_py_random_sample")
    return out, tuple(st)


def _candidates_pinned(population: int, k: int, pin: bool = True) -> torch.Tensor:
    """``py_random_sample(population, k, pin)``, served from the speculative draw when it is valid."""
    global _PREFETCH, _PREFETCH_POOL
    if not PREFETCH_CANDIDATES:
        return py_random_sample(population, k, pin=pin)
    version, internal, gauss = random.getstate()
    out = None
    if _PREFETCH is not None:
        p_pop, p_k, p_from, fut = _PREFETCH
        _PREFETCH = None
        res, after = fut.result()
        if p_pop == population and p_k == k and p_from == internal:
            out, internal = res, after
            random.setstate((version, internal, gauss))
    if out is None:
        out, internal = _draw(internal, population, k, torch.empty(k, dtype=torch.int64, pin_memory=bool(pin)))
        random.setstate((version, internal, gauss))
    if _PREFETCH_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _PREFETCH_POOL = ThreadPoolExecutor(max_workers=1, thread_name_prefix="llp-candidates")
    nxt = torch.empty(k, dtype=torch.int64, pin_memory=bool(pin))   # allocated here: the worker never touches the CUDA context
    _PREFETCH = (population, k, internal, _PREFETCH_POOL.submit(_draw, internal, population, k, nxt))
    return out


_EDGE_ID_CACHE: Dict[tuple, tuple] = {}


def _sorted_edge_ids(edge_index: torch.Tensor, idx: torch.Tensor, num_nodes: int, force_undirected: bool) -> torch.Tensor:
    """Sorted linearised ids of the existing edges, kept per (edge_index tensor, version): the training loops pass the
    same graph every step."""
    import weakref
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_nodes), bool(force_undirected))
    hit = _EDGE_ID_CACHE.get(key)
    if hit is not None and hit[0]() is edge_index:
        return hit[1]
    while len(_EDGE_ID_CACHE) >= 8:
        _EDGE_ID_CACHE.pop(next(iter(_EDGE_ID_CACHE)))
    srt = torch.sort(idx).values
    _EDGE_ID_CACHE[key] = (weakref.ref(edge_index), srt)
    return srt


def _dense_ids_cached(edge_index: torch.Tensor, num_nodes: int):
    """(sorted linearised ids of the non-self-loop edges, their count) of a graph the training loop passes every step:
    computed once per (edge_index tensor, version) — the per-step call then launches nothing for the graph side."""
    import weakref
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_nodes), "dense")
    hit = _EDGE_ID_CACHE.get(key)
    if hit is not None and hit[0]() is edge_index:
        return hit[1]
    while len(_EDGE_ID_CACHE) >= 8:
        _EDGE_ID_CACHE.pop(next(iter(_EDGE_ID_CACHE)))
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    row, col = row[keep], col[keep].clone()
    col[row < col] -= 1
    srt = torch.sort(row * (num_nodes - 1) + col).values
    torch.cuda.current_stream(edge_index.device).synchronize()   # the side stream reads it from the next call on
    val = (srt, int(srt.numel()))
    _EDGE_ID_CACHE[key] = (weakref.ref(edge_index), val)
    return val


_COUNT_PINNED: Dict[int, torch.Tensor] = {}


def _dense_round_on_side_stream(population: int, sample_size: int, taken: torch.Tensor, num_nodes: int, num_neg: int, dev):
    """One round of upstream's dense loop: ``rnd = sample(population, k); rnd = rnd[mask[rnd]]`` cut to ``num_neg`` and
    de-linearised, on a side stream.  Returns ``(edges [2, num_neg], kept ids [num_neg], kept count)``; ``edges`` is valid
    on the CURRENT stream when the call returns and complete only if ``count >= num_neg``."""
    from . import _native as N
    lib = N.require_gpu()
    cur, side = torch.cuda.current_stream(dev), ops._side_stream(dev, "neg")
    rnd_host = _candidates_pinned(population, sample_size)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    cnt_host = _COUNT_PINNED.get(idx)
    if cnt_host is None:
        cnt_host = _COUNT_PINNED[idx] = torch.zeros(1, dtype=torch.int32).pin_memory()
    with torch.cuda.stream(side):
        rnd = rnd_host.to(dev, non_blocking=True)
        edges = torch.empty((2, num_neg), dtype=torch.int64, device=dev)
        kept = torch.empty(num_neg, dtype=torch.int64, device=dev)
        cnt = torch.empty(1, dtype=torch.int32, device=dev)
        nbytes = lib.llp_negative_filter_workspace_bytes(sample_size)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        N.check(lib.llp_negative_filter(rnd.data_ptr(), sample_size, taken.data_ptr(), taken.numel(), num_nodes, num_neg,
                                        kept.data_ptr(), edges.data_ptr(), cnt.data_ptr(), ws.data_ptr(), nbytes,
                                        side.cuda_stream), "llp_negative_filter")
        cnt_host.copy_(cnt, non_blocking=True)
        done = torch.cuda.Event()
        done.record(side)
    done.synchronize()            # waits for the side stream only: the previous training step keeps running
    cur.wait_event(done)
    edges.record_stream(cur)
    kept.record_stream(cur)
    return edges, kept, int(cnt_host[0])


def _not_in_sorted(values: torch.Tensor, sorted_ids: torch.Tensor) -> torch.Tensor:
    """``~isin(values, sorted_ids)`` by binary search — what indexing PyG's N*N - N boolean mask answers, without the mask."""
    if sorted_ids.numel() == 0:
        return torch.ones_like(values, dtype=torch.bool)
    pos = torch.searchsorted(sorted_ids, values).clamp_(max=sorted_ids.numel() - 1)
    return sorted_ids[pos] != values


def negative_sampling(edge_index: torch.Tensor, num_nodes: Optional[int] = None, num_neg_samples: Optional[int] = None,
                      method: str = "sparse", force_undirected: bool = False) -> torch.Tensor:
    """PyG 2.2.0 ``negative_sampling`` for a single (non-bipartite) node set [3P, restated; parity unpinned].
    The candidate ids come from Python's ``random.sample`` on the host exactly as upstream (bit-exact
    indices require the CPython MT19937 stream).  ``method='dense'`` (the training loops, train_teacher_gnn.py:50-51)
    keeps the N*N-N validity mask and the filtering on ``edge_index.device``; ``method='sparse'`` (the split generators,
    utils.py:70-72 / RandomLinkSplit) filters with ``isin`` on the host.  ``force_undirected`` (production split,
    generate_production_split.py:47) samples from the strict upper triangle and returns both directions."""
    if method not in ("dense", "sparse"):
        raise ValueError(method)
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1
    if isinstance(num_nodes, (tuple, list)):
        if num_nodes[0] != num_nodes[1]:
            raise NotImplementedError("bipartite negative sampling is never used by the reference")
        num_nodes = int(num_nodes[0])
    dev = edge_index.device
    fast = method == "dense" and not force_undirected and dev.type == "cuda" and num_nodes >= 2
    cached = _dense_ids_cached(edge_index, num_nodes) if fast else None
    if cached is not None:
        taken, n_existing = cached
        population = num_nodes * num_nodes - num_nodes
    else:
        row, col = edge_index[0].clone(), edge_index[1].clone()
        # edge_index_to_vector
        if force_undirected:
            keep = row < col
            row, col = row[keep], col[keep]
            offset = torch.arange(1, num_nodes, device=dev).cumsum(0)[row]
            idx = row * num_nodes + col - offset
            population = (num_nodes * (num_nodes + 1)) // 2 - num_nodes
        else:
            keep = row != col
            row, col = row[keep], col[keep]
            col[row < col] -= 1
            idx = row * (num_nodes - 1) + col
            population = num_nodes * num_nodes - num_nodes
        n_existing = idx.numel()
        taken = _sorted_edge_ids(edge_index, idx, num_nodes, force_undirected) if method == "dense" else None
    if n_existing >= population:
        return edge_index.new_empty((2, 0))
    if num_neg_samples is None:
        num_neg_samples = edge_index.size(1)
    if force_undirected:
        num_neg_samples = num_neg_samples // 2
    prob = 1.0 - n_existing / population
    sample_size = int(1.1 * num_neg_samples / prob)
    neg_idx = None
    if method == "dense":
        # PyG builds a boolean mask over the whole N*N - N population (1.2 GB per step at the Coauthor-Physics size) and
        # indexes it with the candidates; membership in the SORTED edge ids answers the same question (same candidates
        # kept, same order), on the device, without the mask
        first_round = 0
        if fast and population > sample_size and num_neg_samples > 0:
            # round 1 without a device-wide sync: candidates -> pinned memory -> side stream -> one fused filter /
            # compaction / de-linearisation call; the host only waits for that call's kept count
            edges, kept, count = _dense_round_on_side_stream(population, sample_size, taken, num_nodes, num_neg_samples, dev)
            if count >= num_neg_samples:
                return edges
            neg_idx, first_round = kept[:count], 1   # (rare) continue with upstream's rounds 2 and 3
        for _ in range(first_round, 3):
            rnd = _sample_ids(population, sample_size, dev)
            keep = _not_in_sorted(rnd, taken)
            if neg_idx is not None:   # upstream: mask[neg_idx] = False after an incomplete round
                keep &= _not_in_sorted(rnd, torch.sort(neg_idx).values)
            rnd = rnd[keep]
            neg_idx = rnd if neg_idx is None else torch.cat([neg_idx, rnd])
            if neg_idx.numel() >= num_neg_samples:
                neg_idx = neg_idx[:num_neg_samples]
                break
    else:
        idx_host = idx.cpu()
        for _ in range(3):
            rnd = _sample_ids(population, sample_size, "cpu")
            bad = torch.isin(rnd, idx_host)
            if neg_idx is not None:
                bad |= torch.isin(rnd, neg_idx.cpu())
            rnd = rnd[~bad].to(dev)
            neg_idx = rnd if neg_idx is None else torch.cat([neg_idx, rnd])
            if neg_idx.numel() >= num_neg_samples:
                neg_idx = neg_idx[:num_neg_samples]
                break
    # vector_to_edge_index
    if force_undirected:
        offset = torch.arange(1, num_nodes, device=dev).cumsum(0)
        end = torch.arange(num_nodes, num_nodes * num_nodes, num_nodes, device=dev)
        r = torch.bucketize(neg_idx, end - offset, right=True)
        c = (offset[r] + neg_idx) % num_nodes
        return torch.stack([torch.cat([r, c]), torch.cat([c, r])], dim=0)
    r = neg_idx.div(num_nodes - 1, rounding_mode="floor")
    c = neg_idx % (num_nodes - 1)
    c[r <= c] += 1
    return torch.stack([r, c], dim=0)


_RAND_ON_HOST = False


def draw_rand_on_host(flag: bool) -> None:
    """Parity-test switch: draw the walk/negative random numbers with torch's CPU generator (the stream the CPU
    oracle consumes) instead of the CUDA generator."""
    global _RAND_ON_HOST
    _RAND_ON_HOST = bool(flag)


def random_walk(row: torch.Tensor, col: torch.Tensor, start: torch.Tensor, walk_length: int, p: float = 1, q: float = 1,
                coalesced: bool = True, num_nodes: Optional[int] = None, return_edge_indices: bool = False,
                rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch_cluster 1.6.0 ``random_walk`` (uniform: p == q == 1).  ``rowptr = cumsum(bincount(row))`` and ``col`` is
    used in the given order when ``coalesced=False``.  ``rand`` optionally supplies the ``[B, L]``
    uniform numbers (otherwise ``torch.rand`` on the start tensor's device, as upstream)."""
    if p != 1 or q != 1 or return_edge_indices:
        raise NotImplementedError("only uniform walks returning node sequences are used by the reference")
    dev = start.device
    if num_nodes is None:   # upstream's default: three blocking reductions; pass num_nodes to stay asynchronous
        num_nodes = max(int(row.max()), int(col.max()), int(start.max())) + 1
    if coalesced:
        perm = torch.argsort(row * num_nodes + col)
        row, col = row[perm], col[perm]
        rowptr = _walk_rowptr(row, num_nodes, cache=False)
    else:
        rowptr = _walk_rowptr(row, num_nodes, cache=True)
    if rand is None:
        if _RAND_ON_HOST:
            rand = torch.rand(start.size(0), walk_length).to(dev)
        else:
            rand = torch.rand(start.size(0), walk_length, device=dev)
    return ops.random_walk_with_rand(rowptr, col, start, rand.to(torch.float32))


_ROWPTR_CACHE: Dict[tuple, tuple] = {}


def _walk_rowptr(row: torch.Tensor, num_nodes: int, cache: bool) -> torch.Tensor:
    """``rowptr = cumsum(bincount(row))`` of torch_cluster's walk.  The training loops walk the same
    graph every step (three walks per step): kept per (row tensor, version, num_nodes)."""
    import weakref
    key = (row.data_ptr(), int(row.numel()), row._version, int(num_nodes))
    if cache:
        hit = _ROWPTR_CACHE.get(key)
        if hit is not None and hit[0]() is row:
            return hit[1]
    deg = row.new_zeros(num_nodes)
    deg.scatter_add_(0, row, torch.ones_like(row))
    rowptr = row.new_zeros(num_nodes + 1)
    torch.cumsum(deg, 0, out=rowptr[1:])
    if cache:
        while len(_ROWPTR_CACHE) >= 8:
            _ROWPTR_CACHE.pop(next(iter(_ROWPTR_CACHE)))
        _ROWPTR_CACHE[key] = (weakref.ref(row), rowptr)
    return rowptr


_KS_INDEX: Dict[tuple, torch.Tensor] = {}


def _ks_index(Ks: Sequence[int], dev) -> torch.Tensor:
    """Device tensor of ``K-1`` for every K (cached: building it is a blocking pageable H2D copy)."""
    key = (tuple(Ks), str(dev))
    t = _KS_INDEX.get(key)
    if t is None:
        t = _KS_INDEX[key] = torch.tensor([k - 1 for k in Ks], dtype=torch.int64, device=dev)
    return t


def hits_counts(y_pred_pos: torch.Tensor, y_pred_neg: torch.Tensor, Ks: Sequence[int], group=None, topk_fn=None,
                count_fn=None):
    """Integer hit counts for every K in ONE pass over the scores: returns ``(counts int64 [len(Ks)], n_pos int64 [1])``
    as device tensors (nothing here synchronises with the host on a single rank).
    With ``group`` (a torch.distributed process group) positives and negatives are rank-local shards: each rank
    contributes its top-K_max negatives (all-gather of W*K_max floats), thresholds are taken from the merged
    candidates and the counts are all-reduced — exact and independent of the sharding.
    ``topk_fn`` / ``count_fn`` default to the CUDA kernels (``llp_topk_desc`` / ``llp_count_greater``); they are only
    injectable so the exchange logic can be exercised by the world_size-2 gloo tests on a CPU-only box."""
    topk_fn = topk_fn or ops.topk_desc
    count_fn = count_fn or ops.count_greater
    kmax = max(Ks)
    cand = topk_fn(y_pred_neg, kmax)
    dev = cand.device
    n_pos_total = torch.full((1,), y_pred_pos.numel(), dtype=torch.int64, device=dev)
    n_neg_host = y_pred_neg.numel()
    if group is not None:
        import torch.distributed as dist
        world = dist.get_world_size(group)
        gathered = [torch.empty_like(cand) for _ in range(world)]
        dist.all_gather(gathered, cand, group=group)
        cand = topk_fn(torch.cat(gathered), kmax)
        n_neg = torch.full((1,), n_neg_host, dtype=torch.int64, device=dev)
        dist.all_reduce(n_neg, group=group)
        dist.all_reduce(n_pos_total, group=group)
    # fewer negatives than K  =>  every positive is a hit (ogb: `if len(y_pred_neg) < K: return 1.0`); the top-k
    # list is padded with -inf in that case, which a strict '>' already treats as "always hit" for finite scores.
    thr = cand.index_select(0, _ks_index(Ks, dev))
    counts = count_fn(y_pred_pos, thr)
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(counts, group=group)
        short = n_neg < (_ks_index(Ks, dev) + 1)
        return torch.where(short, n_pos_total.expand_as(counts), counts), n_pos_total
    if n_neg_host < kmax:  # rare: decided on the host (the negative count is a host integer on a single rank)
        short = torch.tensor([n_neg_host < k for k in Ks], device=dev)
        counts = torch.where(short, n_pos_total.expand_as(counts), counts)
    return counts, n_pos_total


def roc_auc_score_device(y_pred_pos: torch.Tensor, y_pred_neg: torch.Tensor, group=None, pairs_fn=None) -> float:
    """``sklearn.metrics.roc_auc_score(cat(ones, zeros), cat(pos, neg))`` of train_teacher_gnn.py:147-153,251-266
    without copying the scores to the host: ``llp_auc_pairs`` counts the (positive, negative) pairs ordered correctly
    and the tied pairs as integers; the only floating-point operation is the final division (in double).
    With ``group`` the scores are rank-local shards: the negatives are all-gathered (ragged shards travel padded and are
    trimmed by the gathered sizes), every rank counts its own positives against all of them and the two
    counters are all-reduced — exact and independent of the sharding.  ``pairs_fn`` is injectable for the gloo tests."""
    pairs_fn = pairs_fn or ops.auc_pairs
    dev = y_pred_pos.device
    neg = y_pred_neg.float().reshape(-1)
    pos = y_pred_pos.float().reshape(-1)
    n_host = (pos.numel(), neg.numel())
    if group is not None:
        n = torch.tensor(n_host, dtype=torch.int64, device=dev)
        import torch.distributed as dist
        world = dist.get_world_size(group)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, n[1:2].clone(), group=group)
        cap = max(int(max(s.item() for s in sizes)), 1)
        padded = torch.full((cap,), float("inf"), dtype=torch.float32, device=dev)
        padded[:neg.numel()] = neg
        gathered = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(gathered, padded, group=group)
        neg = torch.cat([g[:int(s.item())] for g, s in zip(gathered, sizes)])
        dist.all_reduce(n, group=group)
    pairs = pairs_fn(pos, neg)
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(pairs, group=group)
    n_pos, n_neg = (int(t) for t in n.tolist()) if group is not None else n_host
    if n_pos == 0 or n_neg == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    less, equal = (int(t) for t in pairs.tolist())
    return (2 * less + equal) / (2.0 * n_pos * n_neg)


class Evaluator:
    """``ogb.linkproppred.Evaluator`` for the hits@K metric family with mutable ``K``
    (``evaluator.K = K`` at train_teacher_gnn.py:121).  ``eval`` takes the ogb input dict."""

    def __init__(self, name: str = "ogbl-ddi"):
        self.name = name
        self.K = 20
        self.eval_metric = "hits@20"

    def eval(self, input_dict: Dict[str, torch.Tensor]) -> Dict[str, float]:
        if "y_pred_pos" not in input_dict or "y_pred_neg" not in input_dict:
            raise RuntimeError("Missing key of y_pred_pos or y_pred_neg")
        pos, neg = input_dict["y_pred_pos"], input_dict["y_pred_neg"]
        if len(neg) < self.K:
            return {f"hits@{self.K}": 1.0}
        counts, n_pos = hits_counts(pos, neg, [self.K])
        return {f"hits@{self.K}": float(counts[0].item()) / len(pos)}
