"""Synthetic graphs of the BASELINE.json shapes and the split containers the step functions read.

There is no network in this environment, so the Planetoid / Coauthor / OGB downloads of
``src/utils.py:30-50`` cannot run; the drivers and ``bench.py`` use these shape-faithful generators instead
(``data: synthetic`` in every reported number).  Real splits saved by the reference (``../data/<ds>.pkl``,
``train_teacher_gnn.py:310-314``) are plain dicts of ``[E,2]`` tensors and can be passed straight to the step
functions.  Everything here is host-side preparation, outside the hot path.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .shims import Data

SHAPES = {
    # name: (nodes, undirected pairs, feature dim, power_law, feature density or None for gaussian)
    "cora": (2708, 5278, 1433, False, 18 / 1433),
    "citeseer": (3327, 4552, 3703, False, 32 / 3703),
    "pubmed": (19717, 44324, 500, False, 0.1),
    "coauthor-cs": (18333, 81894, 6805, True, 0.0088),
    "coauthor-physics": (34493, 247962, 8415, True, 0.004),
    "amazon-computers": (13752, 245861, 767, True, 0.35),
    "amazon-photos": (7650, 119081, 745, True, 0.35),
    "collab": (235868, 1179052, 128, True, None),
    # BASELINE.json configs[4]: synthetic power-law graph, 10M nodes / 100M undirected pairs (200M messages), 256-d
    "powerlaw-10m": (10_000_000, 100_000_000, 256, True, None),
}


def undirected_graph(num_nodes: int, num_pairs: int, seed: int = 0, power_law: bool = False, exponent: float = 2.5,
                     unique: bool = True, device="cpu") -> torch.Tensor:
    """``[2, 2*pairs]`` symmetrised edge_index sorted by (row, col), no self loops.  ``power_law`` draws endpoints
    with Chung-Lu weights ``rank^(-1/(exponent-1))``.  ``unique=False`` keeps multi-edges (ogbl-collab has them).
    ``device``: where the generator runs (the 10M-node graph of BASELINE.json configs[4] takes minutes on host cores and
    seconds on a GPU; the streams differ between devices, the shapes do not)."""
    g = torch.Generator(device=device).manual_seed(seed)
    m = int(num_pairs * 1.2) + 64

    def draw(k):
        if power_law:
            w = torch.arange(1, num_nodes + 1, dtype=torch.float64, device=device) ** (-1.0 / (exponent - 1.0))
            w = w[torch.randperm(num_nodes, generator=g, device=device)]
            if torch.device(device).type == "cpu":
                return (torch.multinomial(w, k, replacement=True, generator=g),
                        torch.multinomial(w, k, replacement=True, generator=g))
            cdf = torch.cumsum(w / w.sum(), 0)   # inverse-CDF sampling: no limit on the number of draws
            pick = lambda: torch.searchsorted(cdf, torch.rand(k, dtype=torch.float64, generator=g, device=device)).clamp_(max=num_nodes - 1)
            return pick(), pick()
        return (torch.randint(0, num_nodes, (k,), generator=g, device=device),
                torch.randint(0, num_nodes, (k,), generator=g, device=device))

    a, b = draw(m)
    keep = a != b
    lo, hi = torch.minimum(a[keep], b[keep]), torch.maximum(a[keep], b[keep])
    key = lo * num_nodes + hi
    if unique:
        key = torch.unique(key)
    key = key[torch.randperm(key.numel(), generator=g, device=device)[:num_pairs]]
    lo, hi = key // num_nodes, key % num_nodes
    row, col = torch.cat([lo, hi]), torch.cat([hi, lo])
    order = torch.argsort(row * num_nodes + col, stable=True)
    return torch.stack([row[order], col[order]], dim=0)


def features(num_nodes: int, dim: int, density, seed: int = 0, device="cpu") -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed + 17)
    if density is None:
        return torch.randn(num_nodes, dim, generator=g, device=device) * 0.4
    return (torch.rand(num_nodes, dim, generator=g, device=device) < density).float()


def transductive_split(edge_index: torch.Tensor, num_nodes: int, val_ratio: float = 0.05, test_ratio: float = 0.1,
                       seed: int = 234, n_neg_eval: int = 0, device="cpu") -> Dict[str, Dict[str, torch.Tensor]]:
    """Shape of ``do_edge_split`` (src/utils.py:62-105): undirected pairs split train/valid/test, the training
    pairs symmetrised and sorted (``to_undirected``), equal-sized random negatives for valid/test (or
    ``n_neg_eval`` shared-size negatives, the ogbl-collab layout)."""
    g = torch.Generator(device=device).manual_seed(seed)
    und = edge_index[:, edge_index[0] < edge_index[1]].t()
    perm = torch.randperm(und.size(0), generator=g, device=device)
    n_v, n_t = int(val_ratio * und.size(0)), int(test_ratio * und.size(0))
    val, test, train = und[perm[:n_v]], und[perm[n_v:n_v + n_t]], und[perm[n_v + n_t:]]
    tr = torch.cat([train, train.flip(1)], 0)
    tr = tr[torch.argsort(tr[:, 0] * num_nodes + tr[:, 1], stable=True)]
    nv = n_neg_eval or n_v
    nt = n_neg_eval or n_t
    return {
        "train": {"edge": tr},
        "valid": {"edge": val, "edge_neg": torch.randint(0, num_nodes, (nv, 2), generator=g, device=device)},
        "test": {"edge": test, "edge_neg": torch.randint(0, num_nodes, (nt, 2), generator=g, device=device)},
    }


def synthetic_dataset(name: str, seed: int = 0, scale: float = 1.0, device="cpu") -> Tuple[Data, Dict]:
    """(data, split_edge) of the named shape; ``data.adj_t`` is the dense ``[2,E]`` training edge_index exactly as
    the drivers set it (train_teacher_gnn.py:316-317,331).  ``device`` != cpu generates the graph on that device (another
    random stream, same shapes) and returns host tensors."""
    n, pairs, dim, pl, density = SHAPES[name]
    n, pairs = max(int(n * scale), 16), max(int(pairs * scale), 32)
    if torch.device(device).type != "cpu":
        if name == "collab":
            raise ValueError("device generation is for the large generic shapes")
        ei = undirected_graph(n, pairs, seed, pl, device=device)
        split = transductive_split(ei, n, device=device)
        split = {k: {j: t.cpu() for j, t in v.items()} for k, v in split.items()}
        adj = split["train"]["edge"].t().contiguous()
        x = features(n, dim, density, seed, device=device).cpu()
        del ei
        torch.cuda.empty_cache()
        return Data(x=x, adj_t=adj, edge_index=adj), split
    if name == "collab":
        # ogbl-collab: 1,179,052 training pairs (multi-edges kept), 60,084 / 46,329 eval positives, 100,000 negatives
        ei = undirected_graph(n, pairs, seed, True, unique=False)
        g = torch.Generator().manual_seed(seed + 5)
        und = ei[:, ei[0] < ei[1]].t()
        n_v, n_t, n_neg = max(int(60084 * scale), 8), max(int(46329 * scale), 8), max(int(100000 * scale), 16)
        pick = torch.randint(0, und.size(0), (n_v + n_t,), generator=g)
        split = {
            "train": {"edge": und},  # collab trains on one direction per pair (split_edge['train']['edge'])
            "valid": {"edge": und[pick[:n_v]], "edge_neg": torch.randint(0, n, (n_neg, 2), generator=g)},
            "test": {"edge": und[pick[n_v:]], "edge_neg": torch.randint(0, n, (n_neg, 2), generator=g)},
        }
        data = Data(x=features(n, dim, density, seed), adj_t=ei, edge_index=ei)
        return data, split
    ei = undirected_graph(n, pairs, seed, pl)
    split = transductive_split(ei, n)
    adj = split["train"]["edge"].t().contiguous()
    data = Data(x=features(n, dim, density, seed), adj_t=adj, edge_index=adj)
    return data, split


def synthetic_full_graph(name: str, seed: int = 0, scale: float = 1.0) -> Data:
    """The whole (unsplit) synthetic graph of the named shape as a PyG-style ``Data(x, edge_index)`` — the input of the
    split generators (``splits.do_edge_split`` / ``do_production_edge_split``), standing in for ``get_dataset(...)[0]``
    (``src/utils.py:30-50``)."""
    n, pairs, dim, pl, density = SHAPES[name]
    n, pairs = max(int(n * scale), 16), max(int(pairs * scale), 32)
    return Data(x=features(n, dim, density, seed), edge_index=undirected_graph(n, pairs, seed, pl))
