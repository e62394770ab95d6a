"""CPU oracle for the LLP hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``linkless_link_prediction_b200``) never imports it and has no CPU fallback.

Parity status
-------------
* Reference-owned arithmetic (``src/models.py``, ``kl_loss``, the LLP_R block, the
  train/test step functions) is PINNED: ``tests/golden/make_golden.py`` executes the
  reference's own functions in this container (third-party imports stubbed with the
  restatements below) and the resulting fixtures are checked against this oracle.
* Third-party arithmetic (PyG ``SAGEConv`` / ``negative_sampling``, torch_cluster
  ``random_walk``, ogb ``Evaluator``) is **parity unpinned**: torch_geometric==2.2.0,
  torch_scatter, torch_sparse==0.6.16, torch_cluster==1.6.0 and ogb==1.3.6
  (``/root/reference/requirements.txt:1-9``) are not vendored, not installed and there
  is no network.  Their published algorithms are restated here and anchored on the
  reference's call sites plus first-principles known-answer tests.

Everything is plain torch on CPU tensors (fp32 by default, fp64 on request).
"""
from __future__ import annotations

import itertools
import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.data import DataLoader

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# O2/K1-K3: graph structure + mean aggregation
# --------------------------------------------------------------------------------------
def csr_build(edge_index: Tensor, num_nodes: int, by: str = "dst") -> Tuple[Tensor, Tensor, Tensor]:
    """CSR of the message graph.

    ``edge_index`` is the reference's dense ``[2,E]`` LongTensor (``data.adj_t`` at
    ``train_teacher_gnn.py:317,331``); messages flow ``edge_index[0] -> edge_index[1]``
    (PyG ``source_to_target``).  ``by='dst'`` groups messages by destination
    (rows = destinations, cols = sources: the forward aggregation), ``by='src'`` is the
    transpose used by the backward pass.  Ties keep the original edge order (stable),
    which fixes the floating-point summation order.

    Returns ``(rowptr int32 [N+1], col int32 [E], perm int32 [E])`` with
    ``col = other_endpoint[perm]``.
    """
    src, dst = edge_index[0].long(), edge_index[1].long()
    key, val = (dst, src) if by == "dst" else (src, dst)
    perm = torch.sort(key, stable=True).indices
    counts = torch.bincount(key, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr.to(torch.int32), val[perm].to(torch.int32), perm.to(torch.int32)


def scatter_mean(src_rows: Tensor, index: Tensor, dim_size: int) -> Tensor:
    """torch_scatter ``scatter(reduce='mean')`` [3P]: sum, count.clamp_(1), true_divide."""
    out = torch.zeros(dim_size, src_rows.size(1), dtype=src_rows.dtype)
    out.index_add_(0, index, src_rows)
    count = torch.zeros(dim_size, dtype=src_rows.dtype)
    count.index_add_(0, index, torch.ones_like(index, dtype=src_rows.dtype))
    count.clamp_(min=1)
    return out / count.unsqueeze(-1)


def mean_aggregate(x: Tensor, edge_index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
    """``out[d] = sum_{(s->d)} x[s] / max(deg_in[d], 1)`` — the gather→scatter-mean path PyG
    takes when ``edge_index`` is a Tensor."""
    n = x.size(0) if num_nodes is None else num_nodes
    return scatter_mean(x.index_select(0, edge_index[0]), edge_index[1], n)


def spmm_csr(rowptr: Tensor, col: Tensor, x: Tensor, mean: bool, src_scale: Optional[Tensor] = None) -> Tensor:
    """Row-by-row CSR gather-reduce (what the CUDA kernel computes), summing each row's
    neighbours in CSR order.  ``src_scale`` multiplies every gathered row by a per-source
    scalar (used by the transpose-backward: ``gx[s] = sum_d g[d] / deg[d]``)."""
    n = rowptr.numel() - 1
    rp = rowptr.long()
    deg = rp[1:] - rp[:-1]
    rows = torch.repeat_interleave(torch.arange(n), deg)
    vals = x.index_select(0, col.long())
    if src_scale is not None:
        vals = vals * src_scale.index_select(0, col.long()).unsqueeze(-1).to(vals.dtype)
    out = torch.zeros(n, x.size(1), dtype=x.dtype)
    out.index_add_(0, rows, vals)
    if mean:
        out = out / deg.clamp(min=1).to(x.dtype).unsqueeze(-1)
    return out


# --------------------------------------------------------------------------------------
# O2/O3: SAGE convolutions
# --------------------------------------------------------------------------------------
def edge_incidence_plan(u: Tensor, v: Tensor, num_nodes: int) -> Tuple[Tensor, Tensor]:
    """Incidences of an edge batch grouped by node — what autograd's ``index_put_(accumulate=True)`` walks for the
    backward of the two gathers ``h[u] * h[v]`` in front of ``LinkPredictor`` (``train_teacher_gnn.py:58``,
    ``main.py:186,214``, ``models.py:140``), restated as a stable sort so that the summation order is fixed.

    Incidence ``i < M`` is endpoint ``u[i]`` of edge ``i``; incidence ``i >= M`` is endpoint ``v[i-M]`` of edge
    ``i-M``.  Returns ``(rowptr int32 [N+1], meta int32 [2M,2])``: row ``n`` lists, in increasing ``i``, the pairs
    ``(edge, other endpoint of that edge)`` of the incidences whose node is ``n``.
    """
    u, v = u.reshape(-1).long(), v.reshape(-1).long()
    M = u.numel()
    key = torch.cat([u, v])
    order = torch.sort(key, stable=True).indices
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(key, minlength=num_nodes), 0)
    edge = torch.where(order < M, order, order - M)
    other = torch.where(order < M, v[edge], u[edge]) if M else edge
    return rowptr.to(torch.int32), torch.stack([edge, other], 1).to(torch.int32)


def hadamard_backward(h: Tensor, u: Tensor, v: Tensor, dz: Tensor) -> Tensor:
    """``gh[n] = sum_{u[m]==n} dz[m]*h[v[m]] + sum_{v[m]==n} dz[m]*h[u[m]]`` added in the order of
    ``edge_incidence_plan`` (fp32, one fused multiply-add per incidence and element is NOT assumed: plain mul + add)."""
    rowptr, meta = edge_incidence_plan(u, v, h.size(0))
    gh = torch.zeros_like(h, dtype=torch.float32)
    for n in range(h.size(0)):
        for p in range(int(rowptr[n]), int(rowptr[n + 1])):
            m, o = int(meta[p, 0]), int(meta[p, 1])
            gh[n] += dz[m].float() * h[o].float()
    return gh


class SAGEConv(nn.Module):
    """PyG 2.2.0 ``SAGEConv(aggr='mean', root_weight=True, bias=True)`` [3P]:
    ``lin_l(mean_{s->d} x[s]) + lin_r(x)``; ``lin_l`` has the bias, ``lin_r`` none.
    Parameter creation order lin_l.weight, lin_l.bias, lin_r.weight (U(±1/sqrt(fan_in)),
    identical to ``nn.Linear``'s default init)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        out = self.lin_l(mean_aggregate(x, edge_index))
        return out + self.lin_r(x)


class SAGEConvUpdated(nn.Module):
    """``SAGEConv_updated`` (``sageconv_updated.py:65-81``): transform first, then aggregate:
    ``mean_{s->d}(W_l x[s] + b_l) + W_r x``.  Isolated destinations lose ``b_l``."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        out = mean_aggregate(self.lin_l(x), edge_index)
        return out + self.lin_r(x)


# --------------------------------------------------------------------------------------
# O1/O4/O5: model library (``src/models.py``)
# --------------------------------------------------------------------------------------
class SAGE(nn.Module):
    """``models.py:82-119`` (norm_type is always "none" in the drivers)."""

    def __init__(self, data_name, in_channels, hidden_channels, out_channels, num_layers, dropout,
                 conv_layer=SAGEConv, norm_type="none"):
        super().__init__()
        dims = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        self.convs = nn.ModuleList(conv_layer(dims[i], dims[i + 1]) for i in range(num_layers))
        self.norms = nn.ModuleList()
        self.norm_type = norm_type
        self.dropout = dropout

    def reset_parameters(self):
        for c in self.convs:
            c.reset_parameters()

    def forward(self, x, adj_t):
        for conv in self.convs[:-1]:
            x = F.dropout(F.relu(conv(x, adj_t)), p=self.dropout, training=self.training)
        return self.convs[-1](x, adj_t)


class MLP(nn.Module):
    """``models.py:6-54``."""

    def __init__(self, num_layers, input_dim, hidden_dim, output_dim, dropout_ratio, norm_type="none"):
        super().__init__()
        self.num_layers = num_layers
        self.norm_type = norm_type
        self.dropout = nn.Dropout(dropout_ratio)
        self.norms = nn.ModuleList()
        if num_layers == 1:
            dims = [input_dim, output_dim]
        else:
            dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1))

    def reset_parameters(self):
        for layer in self.layers:
            layer.reset_parameters()

    def forward(self, feats):
        h = feats
        for l, layer in enumerate(self.layers):
            h = layer(h)
            if l != self.num_layers - 1:
                h = self.dropout(F.relu(h))
        return h


class LinkPredictor(nn.Module):
    """``models.py:121-150``: ``sigmoid(MLP(x_i * x_j))`` or ``sigmoid(sum(x_i * x_j))``."""

    def __init__(self, predictor, in_channels, hidden_channels, out_channels, num_layers, dropout):
        super().__init__()
        self.predictor = predictor
        dims = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        self.lins = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers))
        self.dropout = dropout

    def reset_parameters(self):
        for lin in self.lins:
            lin.reset_parameters()

    def forward(self, x_i, x_j):
        x = x_i * x_j
        if self.predictor == "mlp":
            for lin in self.lins[:-1]:
                x = F.dropout(F.relu(lin(x)), p=self.dropout, training=self.training)
            x = self.lins[-1](x)
        elif self.predictor == "inner":
            x = torch.sum(x, dim=-1)
        return torch.sigmoid(x)


# --------------------------------------------------------------------------------------
# O7/O10/O11: losses
# --------------------------------------------------------------------------------------
def bce_loss(p: Tensor, y: Tensor) -> Tensor:
    """``nn.BCELoss()`` (mean; logs clamped at -100), ``train_teacher_gnn.py:33,59``."""
    return -(y * torch.clamp(torch.log(p), min=-100.0) + (1 - y) * torch.clamp(torch.log(1 - p), min=-100.0)).mean()


def kl_loss(s: Tensor, t: Tensor, T: float = 1.0) -> Tensor:
    """LLP_D core, ``main.py:27-31``: ``sum softmax(t/T) * (log softmax(t/T) - log_softmax(s/T)) * T^2 / B``."""
    y_s = F.log_softmax(s / T, dim=-1)
    y_t = F.softmax(t / T, dim=-1)
    return F.kl_div(y_s, y_t, reduction="sum") * (T ** 2) / y_s.size(0)


def cosine_loss(s: Tensor, t: Tensor) -> Tensor:
    """``main.py:24-25``."""
    return 1 - F.cosine_similarity(s, t.detach(), dim=-1).mean()


def llp_r_loss(s_r: Tensor, t_r: Tensor, margin: float) -> Tensor:
    """LLP_R block, ``main.py:190-203``.  ``s_r, t_r`` are ``[B_n, K]`` sigmoid scores.
    For every pair i<j (``itertools.combinations`` order): ``y=+1 if t_i > t_j+m``, ``-1 if
    t_i < t_j-m`` else 0; loss = mean over ``B_n*P`` of ``max(0, -y*(s_i-s_j)+m)``.  ``y=0``
    pairs contribute the constant ``margin``."""
    K = s_r.size(1)
    pairs = np.array(list(itertools.combinations(range(K), 2))).T
    i, j = torch.as_tensor(pairs[0]), torch.as_tensor(pairs[1])
    ti, tj = t_r[:, i], t_r[:, j]
    y = torch.zeros_like(ti)
    y[ti > tj + margin] = 1
    y[ti < tj - margin] = -1
    return torch.clamp(-y * (s_r[:, i] - s_r[:, j]) + margin, min=0).mean()


# --------------------------------------------------------------------------------------
# O8: negative sampling (PyG 2.2.0, method='dense') [3P]
# --------------------------------------------------------------------------------------
def negative_sampling_dense(edge_index: Tensor, num_nodes: int, num_neg_samples: int) -> Tensor:
    """Uniform non-edges through the linearised id ``row*(N-1)+col'`` with a bool mask of
    size ``N*N-N`` and up to three rounds of ``random.sample`` (host Python RNG)."""
    row, col = edge_index[0].clone(), edge_index[1].clone()
    keep = row != col
    row, col = row[keep], col[keep]
    col[row < col] -= 1
    idx = row * (num_nodes - 1) + col
    population = num_nodes * num_nodes - num_nodes
    if idx.numel() >= population:
        return edge_index.new_empty((2, 0))
    prob = 1.0 - idx.numel() / population
    sample_size = int(1.1 * num_neg_samples / prob)
    mask = torch.ones(population, dtype=torch.bool)
    mask[idx] = False
    neg_idx = None
    for _ in range(3):
        if population <= sample_size:
            rnd = torch.arange(population)
        else:
            rnd = torch.tensor(random.sample(range(population), sample_size))
        rnd = rnd[mask[rnd]]
        neg_idx = rnd if neg_idx is None else torch.cat([neg_idx, rnd])
        if neg_idx.numel() >= num_neg_samples:
            neg_idx = neg_idx[:num_neg_samples]
            break
        mask[neg_idx] = False
    r = neg_idx.div(num_nodes - 1, rounding_mode="floor")
    c = neg_idx % (num_nodes - 1)
    c[r <= c] += 1
    return torch.stack([r, c], dim=0)


# --------------------------------------------------------------------------------------
# N3: torch_geometric 2.2.0 pieces behind the split generators [3P, restated; parity unpinned].
# Written independently of the product's ``splits.py`` / ``shims.py`` (numpy set arithmetic instead of masks and
# ``isin`` on tensors) so that ``tests/golden/make_split_golden.py`` can run the REFERENCE'S split functions over THESE
# and the product has to reproduce the result index for index: two implementations have to agree, not one with itself.
# --------------------------------------------------------------------------------------
def _linearise(edge_index: Tensor, n: int, undirected: bool) -> Tuple[np.ndarray, int]:
    r, c = edge_index[0].numpy().astype(np.int64), edge_index[1].numpy().astype(np.int64)
    if undirected:      # strict upper triangle, rows packed one after another: id = r*n + c - (1 + 2 + ... + (r+1))
        keep = r < c
        r, c = r[keep], c[keep]
        return r * n + c - (r + 1) * (r + 2) // 2, n * (n + 1) // 2 - n
    keep = r != c       # all ordered pairs without the diagonal: the column index skips the diagonal entry
    r, c = r[keep], c[keep]
    return r * (n - 1) + c - (r < c), n * n - n


def _delinearise(idx: np.ndarray, n: int, undirected: bool) -> Tensor:
    if undirected:
        tri = np.cumsum(np.arange(1, n))                      # tri[r] = 1 + ... + (r+1)
        row_end = np.arange(n, n * n, n) - tri               # first id that belongs to a later row
        r = np.searchsorted(row_end, idx, side="right")
        c = (tri[r] + idx) % n
        return torch.from_numpy(np.stack([np.concatenate([r, c]), np.concatenate([c, r])]))
    r, c = idx // (n - 1), idx % (n - 1)
    c = c + (r <= c)
    return torch.from_numpy(np.stack([r, c]))


def negative_sampling(edge_index: Tensor, num_nodes=None, num_neg_samples: Optional[int] = None, method: str = "sparse",
                      force_undirected: bool = False) -> Tensor:
    """PyG 2.2.0 ``utils.negative_sampling`` (one node set).  Up to three rounds of ``random.sample`` candidates (CPython
    stream), each filtered against the existing edges and the negatives kept so far; 'dense' and 'sparse' only differ
    upstream in HOW they filter (bool mask vs ``isin``), not in what survives."""
    if isinstance(num_nodes, (tuple, list)):
        num_nodes = num_nodes[0]
    n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)
    taken, population = _linearise(edge_index, n, force_undirected)
    if taken.size >= population:
        return edge_index.new_empty((2, 0))
    want = edge_index.size(1) if num_neg_samples is None else int(num_neg_samples)
    if force_undirected:
        want //= 2
    k = int(1.1 * want / (1.0 - taken.size / population))
    forbidden = set(taken.tolist())
    kept: List[int] = []
    for _ in range(3):
        cand = list(range(population)) if population <= k else random.sample(range(population), k)
        # upstream filters a round against the edges and the negatives of EARLIER rounds only: duplicates inside one
        # round cannot occur (sampling without replacement)
        kept.extend(v for v in cand if v not in forbidden)
        if len(kept) >= want:
            kept = kept[:want]
            break
        forbidden.update(kept)
    return _delinearise(np.asarray(kept, dtype=np.int64), n, force_undirected)


class GraphData:
    """Attribute bag standing in for ``torch_geometric.data.Data`` in the split generators."""

    def __init__(self, x=None, edge_index=None, **kw):
        if x is not None:
            self.x = x
        if edge_index is not None:
            self.edge_index = edge_index
        self.__dict__.update(kw)

    @property
    def num_nodes(self):
        return self.x.size(0)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


def add_self_loops(edge_index: Tensor, *_, num_nodes: Optional[int] = None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    loops = torch.arange(n, dtype=edge_index.dtype)
    return torch.cat([edge_index, torch.stack([loops, loops])], dim=1), None


def subgraph(subset: Tensor, edge_index: Tensor, edge_attr=None, relabel_nodes: bool = False, num_nodes=None):
    mask = subset if subset.dtype == torch.bool else torch.zeros(int(num_nodes or edge_index.max() + 1), dtype=torch.bool).index_fill_(0, subset, True)
    keep = mask[edge_index[0]] & mask[edge_index[1]]
    ei = edge_index[:, keep]
    if relabel_nodes:
        new_id = torch.cumsum(mask.long(), 0) - 1      # rank of every kept node among the kept nodes
        ei = new_id[ei]
    return ei, None


def train_test_split_edges(data, val_ratio: float = 0.05, test_ratio: float = 0.1):
    """PyG 2.2.0 ``utils.train_test_split_edges``: RNG order = one ``randperm`` over the undirected pairs, one over the
    non-edges of the strict upper triangle (enumerated in row-major order)."""
    n = data.num_nodes
    r, c = data.edge_index
    data.edge_index = None
    up = r < c
    r, c = r[up], c[up]
    n_v, n_t = int(math.floor(val_ratio * r.numel())), int(math.floor(test_ratio * r.numel()))
    order = torch.randperm(r.numel())
    r, c = r[order], c[order]
    data.val_pos_edge_index = torch.stack([r[:n_v], c[:n_v]])
    data.test_pos_edge_index = torch.stack([r[n_v:n_v + n_t], c[n_v:n_v + n_t]])
    tr = torch.stack([r[n_v + n_t:], c[n_v + n_t:]])
    both = torch.cat([tr, tr.flip(0)], dim=1)                               # to_undirected: both directions,
    key = np.unique(both[0].numpy() * n + both[1].numpy())                  # sorted by (row, col), duplicates dropped
    data.train_pos_edge_index = torch.from_numpy(np.stack([key // n, key % n]))
    free = np.triu(np.ones((n, n), dtype=bool), k=1)
    free[r.numpy(), c.numpy()] = False
    fr, fc = np.nonzero(free)                                                # row-major, like Tensor.nonzero()
    pick = torch.randperm(fr.size)[:n_v + n_t].numpy()
    fr, fc = torch.from_numpy(fr[pick]), torch.from_numpy(fc[pick])
    free[fr.numpy(), fc.numpy()] = False
    data.train_neg_adj_mask = torch.from_numpy(free)
    data.val_neg_edge_index = torch.stack([fr[:n_v], fc[:n_v]])
    data.test_neg_edge_index = torch.stack([fr[n_v:n_v + n_t], fc[n_v:n_v + n_t]])
    return data


class RandomNodeSplit:
    """PyG 2.2.0 ``transforms.RandomNodeSplit`` with the default ``split='train_rest'``."""

    def __init__(self, split="train_rest", num_splits=1, num_train_per_class=20, num_val=500, num_test=1000, key="y"):
        assert split == "train_rest" and num_splits == 1
        self.num_val, self.num_test = num_val, num_test

    def __call__(self, data):
        import copy
        out = copy.copy(data)
        n = data.num_nodes
        nv = round(n * self.num_val) if isinstance(self.num_val, float) else self.num_val
        nt = round(n * self.num_test) if isinstance(self.num_test, float) else self.num_test
        order = torch.randperm(n)
        role = torch.zeros(n, dtype=torch.int64)          # 0 train, 1 val, 2 test
        role[order[:nv]] = 1
        role[order[nv:nv + nt]] = 2
        out.train_mask, out.val_mask, out.test_mask = role == 0, role == 1, role == 2
        return out


class RandomLinkSplit:
    """PyG 2.2.0 ``transforms.RandomLinkSplit`` for one homogeneous graph, defaults as the reference leaves them."""

    def __init__(self, num_val=0.1, num_test=0.2, is_undirected=False, key="edge_label", split_labels=False,
                 add_negative_train_samples=True, neg_sampling_ratio=1.0, disjoint_train_ratio=0.0):
        assert not split_labels and not disjoint_train_ratio and key == "edge_label"
        self.num_val, self.num_test, self.is_undirected = num_val, num_test, is_undirected
        self.neg_train, self.ratio = add_negative_train_samples, neg_sampling_ratio

    def __call__(self, data):
        import copy
        ei = data.edge_index
        if self.is_undirected:
            ids = torch.nonzero(ei[0] <= ei[1]).view(-1)
            ids = ids[torch.randperm(ids.numel())]
        else:
            ids = torch.randperm(ei.size(1))
        nv = int(self.num_val * ids.numel()) if isinstance(self.num_val, float) else self.num_val
        nt = int(self.num_test * ids.numel()) if isinstance(self.num_test, float) else self.num_test
        ntr = ids.numel() - nv - nt
        assert ntr > 0
        parts = {"train": ids[:ntr], "val": ids[ntr:ntr + nv], "test": ids[ntr + nv:]}
        message = {"train": ids[:ntr], "val": ids[:ntr], "test": ids[:ntr + nv]}   # edges each split may look at
        n_neg = {"train": int(ntr * self.ratio) if self.neg_train else 0, "val": int(nv * self.ratio), "test": int(nt * self.ratio)}
        total = sum(n_neg.values())
        neg = negative_sampling(ei, (data.num_nodes, data.num_nodes), num_neg_samples=total, method="sparse")
        if neg.size(1) < total:
            scale = neg.size(1) / total
            n_neg["train"], n_neg["val"] = int(n_neg["train"] * scale), int(n_neg["val"] * scale)
            n_neg["test"] = neg.size(1) - n_neg["train"] - n_neg["val"]
        # the negative columns are dealt out val first, then test, then train
        start = {"val": 0, "test": n_neg["val"], "train": n_neg["val"] + n_neg["test"]}
        end = {"val": n_neg["val"], "test": n_neg["val"] + n_neg["test"], "train": neg.size(1)}
        outs = []
        for name in ("train", "val", "test"):
            d = copy.copy(data)
            m = ei[:, message[name]]
            d.edge_index = torch.cat([m, m.flip(0)], dim=1) if self.is_undirected else m
            pos = ei[:, parts[name]]
            mine = neg[:, start[name]:end[name]]
            label = torch.ones(pos.size(1))
            if mine.numel() > 0:
                label = torch.cat([label, torch.zeros(mine.size(1))])
                pos = torch.cat([pos, mine], dim=1)
            d.edge_label, d.edge_label_index = label, pos
            outs.append(d)
        return tuple(outs)


# --------------------------------------------------------------------------------------
# O9: random walks (torch_cluster 1.6.0, uniform, coalesced=False) [3P]
# --------------------------------------------------------------------------------------
def walk_rowptr(row: Tensor, num_nodes: int) -> Tensor:
    """``rowptr = cumsum(bincount(row))``; ``col`` is used in the given order."""
    deg = torch.zeros(num_nodes, dtype=torch.int64)
    deg.scatter_add_(0, row, torch.ones_like(row))
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    torch.cumsum(deg, 0, out=rowptr[1:])
    return rowptr


def random_walk_with_rand(rowptr: Tensor, col: Tensor, start: Tensor, rand: Tensor) -> Tensor:
    """One uniform walk per start.  ``rand`` is the ``[B, L]`` fp32 tensor the kernel draws with
    ``torch.rand``; per hop ``n = col[rowptr[n] + int64(rand * deg)]`` (fp32 multiply, truncate),
    a node without out-edges stays put.  Returns ``[B, L+1]`` int64."""
    B, L = rand.shape
    out = torch.empty(B, L + 1, dtype=torch.int64)
    cur = start.clone().long()
    out[:, 0] = cur
    for l in range(L):
        rs, re = rowptr[cur], rowptr[cur + 1]
        deg = re - rs
        off = (rand[:, l].float() * deg.float()).long()
        e = rs + off
        if col.numel() == 0:  # no edges at all: every walker stays where it is
            out[:, l + 1] = cur
            continue
        nxt = col[torch.where(deg > 0, e, torch.zeros_like(e))]
        cur = torch.where(deg > 0, nxt, cur)
        out[:, l + 1] = cur
    return out


def random_walk(row: Tensor, col: Tensor, start: Tensor, walk_length: int, coalesced: bool = False,
                num_nodes: Optional[int] = None) -> Tensor:
    if num_nodes is None:
        num_nodes = max(int(row.max()), int(col.max()), int(start.max())) + 1
    if coalesced:
        perm = torch.argsort(row * num_nodes + col)
        row, col = row[perm], col[perm]
    rowptr = walk_rowptr(row, num_nodes)
    rand = torch.rand(start.size(0), walk_length)
    return random_walk_with_rand(rowptr, col, start, rand)


def neighbor_samplers(row, col, sample, x, step, ps_method, ns_rate, hops):
    """``main.py:33-50`` (device moves dropped: the oracle is CPU-only)."""
    batch = sample
    if ps_method == "rw":
        pos_batch = random_walk(row, col, batch, walk_length=step * hops, coalesced=False)
    else:
        pos_batch = None
        for _ in range(step):
            w = random_walk(row, col, batch, walk_length=hops, coalesced=False)
            pos_batch = w if pos_batch is None else torch.cat((pos_batch, w[:, 1:]), 1)
    neg_batch = torch.randint(0, x.size(0), (batch.numel(), step * hops * ns_rate), dtype=torch.long)
    return pos_batch, neg_batch


# --------------------------------------------------------------------------------------
# O12: Hits@K (ogb 1.3.6 Evaluator._eval_hits, torch branch) [3P]
# --------------------------------------------------------------------------------------
def hits_at_k(y_pred_pos: Tensor, y_pred_neg: Tensor, K: int) -> float:
    if len(y_pred_neg) < K:
        return 1.0
    kth = torch.topk(y_pred_neg, K)[0][-1]
    return float(torch.sum(y_pred_pos > kth)) / len(y_pred_pos)


def hits_counts(y_pred_pos: Tensor, y_pred_neg: Tensor, Ks: Sequence[int]) -> List[int]:
    """Integer hit counts (``len(pos)`` when ``len(neg) < K``) — what the kernel is compared to bit-exactly."""
    out = []
    for K in Ks:
        if len(y_pred_neg) < K:
            out.append(len(y_pred_pos))
        else:
            kth = np.sort(y_pred_neg.numpy())[-K]
            out.append(int((y_pred_pos.numpy() > kth).sum()))
    return out


# --------------------------------------------------------------------------------------
# N2: ROC-AUC (sklearn.metrics.roc_auc_score, ``train_teacher_gnn.py:147-153,251-266``) [3P]
# --------------------------------------------------------------------------------------
def auc_pairs(y_pred_pos: Tensor, y_pred_neg: Tensor) -> Tuple[int, int]:
    """``(#{(p,n): neg_n < pos_p}, #{(p,n): neg_n == pos_p})`` — the integer content of the binary ROC-AUC.
    sklearn integrates the ROC curve (thresholds at the distinct scores) with the trapezoid rule, which equals the
    Mann-Whitney statistic with ties at half weight; ``tests/test_oracle_golden.py`` pins this against sklearn."""
    pos = y_pred_pos.detach().to(torch.float32).numpy() + np.float32(0.0)
    neg = np.sort(y_pred_neg.detach().to(torch.float32).numpy() + np.float32(0.0))
    lb = np.searchsorted(neg, pos, side="left").astype(np.int64)
    ub = np.searchsorted(neg, pos, side="right").astype(np.int64)
    return int(lb.sum()), int((ub - lb).sum())


def roc_auc(y_pred_pos: Tensor, y_pred_neg: Tensor) -> float:
    if len(y_pred_pos) == 0 or len(y_pred_neg) == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    less, equal = auc_pairs(y_pred_pos, y_pred_neg)
    return (2 * less + equal) / (2.0 * len(y_pred_pos) * len(y_pred_neg))


class Evaluator:
    """Stand-in for ``ogb.linkproppred.Evaluator(name='ogbl-ddi')`` with mutable ``K``
    (``train_teacher_gnn.py:394,120-121``)."""

    def __init__(self, name: str = "ogbl-ddi"):
        self.name, self.K = name, 20

    def eval(self, input_dict: Dict[str, Tensor]) -> Dict[str, float]:
        return {f"hits@{self.K}": hits_at_k(input_dict["y_pred_pos"], input_dict["y_pred_neg"], self.K)}


# --------------------------------------------------------------------------------------
# O13: scoring + O14: optimiser tail
# --------------------------------------------------------------------------------------
@torch.no_grad()
def score_edges(predictor: nn.Module, h: Tensor, edges: Tensor, batch_size: int) -> Tensor:
    """One scoring loop of ``test_transductive`` (``train_teacher_gnn.py:94-98``); ``edges`` is ``[n,2]``."""
    preds = []
    for perm in DataLoader(range(edges.size(0)), batch_size):
        e = edges[perm].t()
        preds += [predictor(h[e[0]], h[e[1]]).squeeze()]
    return torch.cat([p.reshape(-1) for p in preds], dim=0)


def clip_grad_norm(params: Sequence[Tensor], max_norm: float) -> Tensor:
    """``torch.nn.utils.clip_grad_norm_`` semantics: ``coef = min(1, max_norm/(||g||+1e-6))``."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    """``torch.optim.Adam`` single-tensor update (no weight decay, no amsgrad)."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# O6: teacher train step / epoch (``train_teacher_gnn.py:21-73``)
# --------------------------------------------------------------------------------------
def teacher_step(model, predictor, x, adj_t, pos_edge, neg_edge, optimizer, encoder_name="sage") -> float:
    """One mini-batch of the teacher loop with the batch's edges given explicitly
    (``pos_edge``, ``neg_edge`` are ``[2,B]``)."""
    optimizer.zero_grad()
    h = model(x) if encoder_name == "mlp" else model(x, adj_t)
    train_edges = torch.cat((pos_edge, neg_edge), dim=-1)
    label = torch.cat((torch.ones(pos_edge.size(1)), torch.zeros(neg_edge.size(1))), dim=0).to(h.dtype)
    out = predictor(h[train_edges[0]], h[train_edges[1]]).squeeze()
    loss = bce_loss(out, label)
    loss.backward()
    clip_grad_norm(list(model.parameters()), 1.0)
    clip_grad_norm(list(predictor.parameters()), 1.0)
    optimizer.step()
    return loss.item()


def teacher_train_epoch(model, predictor, x, adj_t, pos_train_edge, optimizer, batch_size,
                        encoder_name="sage", dataset="cora") -> float:
    """Full epoch in the reference's RNG order (DataLoader shuffle, then negatives per batch)."""
    row, col = adj_t
    edge_index = torch.stack([col, row], dim=0)
    model.train()
    predictor.train()
    total_loss = total_examples = 0
    for perm in DataLoader(range(pos_train_edge.size(0)), batch_size, shuffle=True):
        edge = pos_train_edge[perm].t()
        if dataset != "collab":
            neg_edge = negative_sampling_dense(edge_index, x.size(0), perm.size(0))
        else:
            neg_edge = torch.randint(0, x.size(0), edge.size(), dtype=torch.long)
        loss = teacher_step(model, predictor, x, adj_t, edge, neg_edge, optimizer, encoder_name)
        total_loss += loss * edge.size(1)
        total_examples += edge.size(1)
    return total_loss / total_examples


@torch.no_grad()
def test_transductive(model, predictor, x, adj_t, split_edge, batch_size, encoder_name="sage",
                      dataset="cora") -> Tuple[Dict[str, Tuple[float, float]], Tensor]:
    """``train_teacher_gnn.py:76-155`` including the AUC row (:147-153)."""
    model.eval()
    predictor.eval()
    h = model(x) if encoder_name == "mlp" else model(x, adj_t)
    pv = score_edges(predictor, h, split_edge["valid"]["edge"], batch_size)
    nv = score_edges(predictor, h, split_edge["valid"]["edge_neg"], batch_size)
    pt = score_edges(predictor, h, split_edge["test"]["edge"], batch_size)
    nt = score_edges(predictor, h, split_edge["test"]["edge_neg"], batch_size)
    results = {}
    for K in ([10, 20, 30, 50] if dataset != "collab" else [10, 50, 100]):
        results[f"Hits@{K}"] = (hits_at_k(pv, nv, K), hits_at_k(pt, nt, K))
    results["AUC"] = (roc_auc(pv, nv), roc_auc(pt, nt))
    return results, h


@torch.no_grad()
def test_production(model, predictor, val_data, inference_data, test_edge_bundle, negative_samples, batch_size,
                    encoder_name="sage"):
    """``train_teacher_gnn.py:157-268``: validation on the training graph, testing on the inference graph (old-old /
    old-new / new-new buckets against one global negative set); ``results[key] = (valid, test, old_old, old_new,
    new_new)``; returns the validation-graph embeddings like the reference (``saved_h``)."""
    model.eval()
    predictor.eval()
    h = model(val_data.x) if encoder_name == "mlp" else model(val_data.x, val_data.edge_index)
    saved_h = h
    val_edges = val_data.edge_label_index.t()
    pos_valid = score_edges(predictor, h, val_edges[val_data.edge_label.bool()], batch_size)
    neg_valid = score_edges(predictor, h, val_edges[(1 - val_data.edge_label).bool()], batch_size)
    h = model(inference_data.x) if encoder_name == "mlp" else model(inference_data.x, inference_data.edge_index)
    old_old, old_new, new_new, test_all = (t.t() for t in test_edge_bundle[:4])
    pos_test, oo, on, nn_ = (score_edges(predictor, h, e, batch_size) for e in (test_all, old_old, old_new, new_new))
    neg_test = score_edges(predictor, h, negative_samples.t(), batch_size)
    pairs = [(pos_valid, neg_valid), (pos_test, neg_test), (oo, neg_test), (on, neg_test), (nn_, neg_test)]
    results = {f"Hits@{K}": tuple(hits_at_k(p, n, K) for p, n in pairs) for K in (10, 20, 30, 50)}
    results["AUC"] = tuple(roc_auc(p, n) for p, n in pairs)
    return results, saved_h


# --------------------------------------------------------------------------------------
# O15: student KD step (``main.py:147-236``, full-batch variant)
# --------------------------------------------------------------------------------------
def student_losses(model, predictor, t_h, teacher_predictor, x, samples, pos_edge, neg_edge, node_perm,
                   margin: float) -> Dict[str, Tensor]:
    """All loss terms of one student step with samples / edges given explicitly.
    ``samples`` is ``[B_n, 1+K]`` (anchor, walk contexts, random contexts)."""
    h = model(x)
    K = samples.size(1) - 1
    anchor = samples[:, 0]
    batch_emb = h[anchor].reshape(-1, 1, h.size(1)).repeat(1, K, 1)
    t_emb = t_h[anchor].reshape(-1, 1, t_h.size(1)).repeat(1, K, 1)
    s_r = predictor(batch_emb, h[samples[:, 1:]]).reshape(samples.size(0), K)
    t_r = teacher_predictor(t_emb, t_h[samples[:, 1:]]).reshape(samples.size(0), K)
    llp_d = kl_loss(s_r, t_r, 1)
    llp_r = llp_r_loss(s_r, t_r, margin)
    train_edges = torch.cat((pos_edge, neg_edge), dim=-1)
    label = torch.cat((torch.ones(pos_edge.size(1)), torch.zeros(neg_edge.size(1))), dim=0).to(h.dtype)
    out = predictor(h[train_edges[0]], h[train_edges[1]]).squeeze()
    label_loss = bce_loss(out, label)
    t_out = teacher_predictor(t_h[train_edges[0]], t_h[train_edges[1]]).squeeze().detach()
    return {
        "label": label_loss, "llp_d": llp_d, "llp_r": llp_r,
        "kd_rm": cosine_loss(h[node_perm], t_h[node_perm]), "kd_lm": F.mse_loss(out, t_out),
        "s_r": s_r, "t_r": t_r, "out": out,
    }


def student_step(model, predictor, t_h, teacher_predictor, x, samples, pos_edge, neg_edge, node_perm,
                 optimizer, True_label=1.0, LLP_D=1.0, LLP_R=1.0, KD_RM=0.0, KD_LM=0.0, margin=0.1) -> float:
    optimizer.zero_grad()
    t = student_losses(model, predictor, t_h, teacher_predictor, x, samples, pos_edge, neg_edge, node_perm, margin)
    loss = True_label * t["label"] + KD_RM * t["kd_rm"] + KD_LM * t["kd_lm"] + LLP_D * t["llp_d"] + LLP_R * t["llp_r"]
    loss.backward()
    clip_grad_norm(list(model.parameters()), 1.0)
    clip_grad_norm(list(predictor.parameters()), 1.0)
    optimizer.step()
    return loss.item()


def student_train_epoch(model, predictor, t_h, teacher_predictor, x, adj_t, pos_train_edge, optimizer, args,
                        dataset: str = "cora") -> float:
    """One epoch of the full-batch student loop in the reference's RNG order (``main.py:147-236``): node loader created
    first (:167), link loader per epoch (:168); per step ``next(node_loader)`` (:171), encoder forward (:173), context
    sampling (:180: ``torch.rand`` per walk, then the CPU ``randint``), negative edges (:205-209), losses, clip x2, Adam.
    Returns ``total_loss / total_examples`` (:232-236)."""
    row, col = adj_t
    edge_index = torch.stack([col, row], dim=0)
    model.train()
    predictor.train()
    total_loss = total_examples = 0
    node_loader = iter(DataLoader(range(x.size(0)), args.node_batch_size, shuffle=True))
    for link_perm in DataLoader(range(pos_train_edge.size(0)), args.link_batch_size, shuffle=True):
        node_perm = next(node_loader)
        edge = pos_train_edge[link_perm].t()
        pos_sample, neg_sample = neighbor_samplers(row, col, node_perm, x, args.rw_step, args.ps_method, args.ns_rate,
                                                   args.hops)
        samples = torch.cat((pos_sample, neg_sample), 1)
        if dataset != "collab":
            neg_edge = negative_sampling_dense(edge_index, x.size(0), link_perm.size(0))
        else:
            neg_edge = torch.randint(0, x.size(0), [edge.size(0), edge.size(1)], dtype=torch.long)
        loss = student_step(model, predictor, t_h, teacher_predictor, x, samples, edge, neg_edge, node_perm, optimizer,
                            True_label=args.True_label, LLP_D=args.LLP_D, LLP_R=args.LLP_R, KD_RM=args.KD_RM,
                            KD_LM=args.KD_LM, margin=args.margin)
        total_loss += loss * edge.size(1)
        total_examples += edge.size(1)
    return total_loss / total_examples


# --------------------------------------------------------------------------------------
# Synthetic graphs of the BASELINE.json shapes — shared by tests and bench
# --------------------------------------------------------------------------------------
def synthetic_undirected_graph(num_nodes: int, num_undirected: int, seed: int = 0, power_law: bool = False,
                               exponent: float = 2.5) -> Tensor:
    """``[2, 2*num_undirected]`` symmetrised, (row, col)-sorted edge_index without self loops.
    Uniform endpoints, or Chung-Lu style power-law endpoint weights when ``power_law``.
    Duplicate pairs are removed (so the count can come out slightly lower)."""
    g = torch.Generator().manual_seed(seed)
    m = int(num_undirected * 1.15) + 16
    if power_law:
        w = (torch.arange(1, num_nodes + 1, dtype=torch.float64)) ** (-1.0 / (exponent - 1.0))
        w = w[torch.randperm(num_nodes, generator=g)]
        a = torch.multinomial(w, m, replacement=True, generator=g)
        b = torch.multinomial(w, m, replacement=True, generator=g)
    else:
        a = torch.randint(0, num_nodes, (m,), generator=g)
        b = torch.randint(0, num_nodes, (m,), generator=g)
    keep = a != b
    a, b = a[keep], b[keep]
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    key = torch.unique(lo * num_nodes + hi)
    key = key[torch.randperm(key.numel(), generator=g)[:num_undirected]]
    lo, hi = key // num_nodes, key % num_nodes
    row, col = torch.cat([lo, hi]), torch.cat([hi, lo])
    order = torch.argsort(row * num_nodes + col)
    return torch.stack([row[order], col[order]], dim=0)
