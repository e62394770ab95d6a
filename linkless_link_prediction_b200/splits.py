"""Split generators of the reference, restated without torch_geometric.  Host-side preparation that
runs once per dataset, outside the hot path: everything here is plain torch on whatever device the inputs live on.

Reference-owned glue (pinned by ``tests/golden/split_golden.pt``, produced by running the reference's own functions over
the third-party restatements below — ``tests/golden/make_split_golden.py``):

* ``do_edge_split``            — ``src/utils.py:62-105`` (== ``src/generate_production_split.py:97-141``)
* ``split_edges``              — ``src/generate_production_split.py:14-30``
* ``do_production_edge_split`` — ``src/generate_production_split.py:32-95``

Third-party pieces [3P] restated from torch_geometric 2.2.0 (``requirements.txt:8``; not installed here, so **parity
unpinned**): ``train_test_split_edges``, ``to_undirected`` / ``coalesce``, ``add_self_loops``, ``subgraph``,
``RandomNodeSplit(split='train_rest')``, ``RandomLinkSplit(is_undirected=True)`` and (in ``shims.py``)
``negative_sampling``.  They consume the torch / CPython RNG streams in upstream's order, so under the same torch version a
seed reproduces upstream's split; the ``.pkl`` containers (``train_teacher_gnn.py:310-314,348,366``) are the same plain
dict / tuple structures, so cached artefacts interchange with the reference.
"""
from __future__ import annotations

import copy
import math
import random
from typing import Optional, Tuple

import torch

from .shims import Data, negative_sampling


# ------------------------------------------------------------------------------------------------
# torch_geometric.utils restatements [3P]
# ------------------------------------------------------------------------------------------------
def coalesce(edge_index: torch.Tensor, num_nodes: Optional[int] = None) -> torch.Tensor:
    """Sort by (row, col) and drop duplicate edges."""
    n = int(edge_index.max()) + 1 if (num_nodes is None and edge_index.numel()) else int(num_nodes or 0)
    key = edge_index[0] * n + edge_index[1]
    key = torch.unique(key, sorted=True)
    return torch.stack([key.div(n, rounding_mode="floor"), key % n], dim=0) if n else edge_index


def to_undirected(edge_index: torch.Tensor, num_nodes: Optional[int] = None) -> torch.Tensor:
    row, col = edge_index[0], edge_index[1]
    both = torch.stack([torch.cat([row, col]), torch.cat([col, row])], dim=0)
    return coalesce(both, num_nodes)


def add_self_loops(edge_index: torch.Tensor, edge_attr=None, fill_value=None, num_nodes: Optional[int] = None):
    n = int(num_nodes) if num_nodes is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1), None


def subgraph(subset: torch.Tensor, edge_index: torch.Tensor, edge_attr=None, relabel_nodes: bool = False,
             num_nodes: Optional[int] = None):
    """Edges whose endpoints both lie in ``subset`` (a bool mask or an index tensor)."""
    dev = edge_index.device
    if subset.dtype == torch.bool:
        n = subset.numel()
        node_mask = subset
    else:
        n = int(num_nodes) if num_nodes is not None else int(edge_index.max()) + 1
        node_mask = torch.zeros(n, dtype=torch.bool, device=dev)
        node_mask[subset] = True
    edge_mask = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    ei = edge_index[:, edge_mask]
    if relabel_nodes:
        node_idx = torch.zeros(n, dtype=torch.long, device=dev)
        node_idx[node_mask] = torch.arange(int(node_mask.sum()), device=dev)
        ei = node_idx[ei]
    return ei, None


def train_test_split_edges(data: Data, val_ratio: float = 0.05, test_ratio: float = 0.1) -> Data:
    """torch_geometric 2.2.0 ``utils.train_test_split_edges`` (deprecated upstream, still what ``do_edge_split`` calls):
    undirected pairs -> val / test / train positives (train symmetrised), negatives drawn from the N x N strict upper
    triangle mask."""
    num_nodes = data.num_nodes
    row, col = data.edge_index
    data.edge_index = None
    mask = row < col
    row, col = row[mask], col[mask]
    n_v = int(math.floor(val_ratio * row.size(0)))
    n_t = int(math.floor(test_ratio * row.size(0)))
    perm = torch.randperm(row.size(0))
    row, col = row[perm], col[perm]
    data.val_pos_edge_index = torch.stack([row[:n_v], col[:n_v]], dim=0)
    data.test_pos_edge_index = torch.stack([row[n_v:n_v + n_t], col[n_v:n_v + n_t]], dim=0)
    data.train_pos_edge_index = to_undirected(torch.stack([row[n_v + n_t:], col[n_v + n_t:]], dim=0))
    neg_adj_mask = torch.ones(num_nodes, num_nodes, dtype=torch.uint8)
    neg_adj_mask = neg_adj_mask.triu(diagonal=1).to(torch.bool)
    neg_adj_mask[row, col] = 0
    neg_row, neg_col = neg_adj_mask.nonzero(as_tuple=False).t()
    perm = torch.randperm(neg_row.size(0))[:n_v + n_t]
    neg_row, neg_col = neg_row[perm], neg_col[perm]
    neg_adj_mask[neg_row, neg_col] = 0
    data.train_neg_adj_mask = neg_adj_mask
    data.val_neg_edge_index = torch.stack([neg_row[:n_v], neg_col[:n_v]], dim=0)
    data.test_neg_edge_index = torch.stack([neg_row[n_v:n_v + n_t], neg_col[n_v:n_v + n_t]], dim=0)
    return data


class RandomNodeSplit:
    """torch_geometric 2.2.0 ``transforms.RandomNodeSplit(split='train_rest')``: ``num_val`` / ``num_test`` nodes (float =
    fraction, rounded) off one ``randperm``, the rest train."""

    def __init__(self, split: str = "train_rest", num_splits: int = 1, num_train_per_class: int = 20, num_val=500,
                 num_test=1000, key: Optional[str] = "y"):
        if split != "train_rest" or num_splits != 1:
            raise NotImplementedError("the reference only uses the default 'train_rest' split")
        self.num_val, self.num_test = num_val, num_test

    def __call__(self, data: Data) -> Data:
        data = copy.copy(data)
        n = data.num_nodes
        num_val = round(n * self.num_val) if isinstance(self.num_val, float) else self.num_val
        num_test = round(n * self.num_test) if isinstance(self.num_test, float) else self.num_test
        train_mask = torch.zeros(n, dtype=torch.bool)
        val_mask = torch.zeros(n, dtype=torch.bool)
        test_mask = torch.zeros(n, dtype=torch.bool)
        perm = torch.randperm(n)
        val_mask[perm[:num_val]] = True
        test_mask[perm[num_val:num_val + num_test]] = True
        train_mask[perm[num_val + num_test:]] = True
        data.train_mask, data.val_mask, data.test_mask = train_mask, val_mask, test_mask
        return data


class RandomLinkSplit:
    """torch_geometric 2.2.0 ``transforms.RandomLinkSplit`` for one homogeneous graph with the defaults the reference
    leaves untouched (``add_negative_train_samples=True``, ``neg_sampling_ratio=1.0``, ``disjoint_train_ratio=0``,
    ``split_labels=False``, ``key='edge_label'``)."""

    def __init__(self, num_val=0.1, num_test=0.2, is_undirected: bool = False, key: str = "edge_label",
                 split_labels: bool = False, add_negative_train_samples: bool = True, neg_sampling_ratio: float = 1.0,
                 disjoint_train_ratio=0.0):
        if split_labels or disjoint_train_ratio or key != "edge_label":
            raise NotImplementedError("only the options the reference uses are restated")
        self.num_val, self.num_test, self.is_undirected = num_val, num_test, is_undirected
        self.add_negative_train_samples, self.neg_sampling_ratio = add_negative_train_samples, neg_sampling_ratio

    def __call__(self, data: Data) -> Tuple[Data, Data, Data]:
        train_data, val_data, test_data = copy.copy(data), copy.copy(data), copy.copy(data)
        edge_index = data.edge_index
        dev = edge_index.device
        if self.is_undirected:
            mask = edge_index[0] <= edge_index[1]
            perm = mask.nonzero(as_tuple=False).view(-1)
            perm = perm[torch.randperm(perm.size(0), device=perm.device)]
        else:
            perm = torch.randperm(edge_index.size(1), device=dev)
        num_val = int(self.num_val * perm.numel()) if isinstance(self.num_val, float) else self.num_val
        num_test = int(self.num_test * perm.numel()) if isinstance(self.num_test, float) else self.num_test
        num_train = perm.numel() - num_val - num_test
        if num_train <= 0:
            raise ValueError("Insufficient number of edges for training")
        train_edges = perm[:num_train]
        val_edges = perm[num_train:num_train + num_val]
        test_edges = perm[num_train + num_val:]
        train_val_edges = perm[:num_train + num_val]

        def split(out: Data, index: torch.Tensor) -> None:
            ei = edge_index[:, index]
            out.edge_index = torch.cat([ei, ei.flip([0])], dim=-1) if self.is_undirected else ei

        split(train_data, train_edges)
        split(val_data, train_edges)
        split(test_data, train_val_edges)
        num_neg_train = int(num_train * self.neg_sampling_ratio) if self.add_negative_train_samples else 0
        num_neg_val = int(num_val * self.neg_sampling_ratio)
        num_neg_test = int(num_test * self.neg_sampling_ratio)
        num_neg = num_neg_train + num_neg_val + num_neg_test
        n = data.num_nodes
        neg_edge_index = negative_sampling(edge_index, (n, n), num_neg_samples=num_neg, method="sparse")
        num_neg_found = neg_edge_index.size(1)          # "adjust ratio if not enough negative edges exist"
        if num_neg_found < num_neg:
            ratio = num_neg_found / num_neg
            num_neg_train = int(num_neg_train * ratio)
            num_neg_val = int(num_neg_val * ratio)
            num_neg_test = num_neg_found - num_neg_train - num_neg_val

        def label(out: Data, index: torch.Tensor, neg: torch.Tensor) -> None:
            ei = edge_index[:, index]
            edge_label = torch.ones(index.numel(), device=dev)
            if neg.numel() > 0:
                edge_label = torch.cat([edge_label, edge_label.new_zeros(neg.size(1))], dim=0)
                ei = torch.cat([ei, neg], dim=-1)
            out.edge_label, out.edge_label_index = edge_label, ei

        label(train_data, train_edges, neg_edge_index[:, num_neg_val + num_neg_test:])
        label(val_data, val_edges, neg_edge_index[:, :num_neg_val])
        label(test_data, test_edges, neg_edge_index[:, num_neg_val:num_neg_val + num_neg_test])
        return train_data, val_data, test_data


# ------------------------------------------------------------------------------------------------
# the reference's own split functions
# ------------------------------------------------------------------------------------------------
def do_edge_split(dataset, fast_split: bool = False, val_ratio: float = 0.05, test_ratio: float = 0.1, split_seed: int = 234):
    """``src/utils.py:62-105``: ``split_edge`` dict of ``[E,2]`` tensors (``train/valid/test`` x ``edge/edge_neg``) —
    the object ``train_teacher_gnn.py:310-314`` caches as ``../data/<ds>.pkl``."""
    data = copy.copy(dataset[0])
    random.seed(split_seed)
    torch.manual_seed(split_seed)
    if not fast_split:
        data = train_test_split_edges(data, val_ratio, test_ratio)
        edge_index, _ = add_self_loops(data.train_pos_edge_index)
        data.train_neg_edge_index = negative_sampling(edge_index, num_nodes=data.num_nodes,
                                                      num_neg_samples=data.train_pos_edge_index.size(1))
    else:
        num_nodes = data.num_nodes
        row, col = data.edge_index
        mask = row < col
        row, col = row[mask], col[mask]
        n_v = int(math.floor(val_ratio * row.size(0)))
        n_t = int(math.floor(test_ratio * row.size(0)))
        perm = torch.randperm(row.size(0))
        row, col = row[perm], col[perm]
        data.val_pos_edge_index = torch.stack([row[:n_v], col[:n_v]], dim=0)
        data.test_pos_edge_index = torch.stack([row[n_v:n_v + n_t], col[n_v:n_v + n_t]], dim=0)
        data.train_pos_edge_index = torch.stack([row[n_v + n_t:], col[n_v + n_t:]], dim=0)
        # negatives: (i, j) and (j, i) may both appear (reference comment, utils.py:91)
        neg = negative_sampling(data.edge_index, num_nodes=num_nodes, num_neg_samples=row.size(0))
        data.val_neg_edge_index = neg[:, :n_v]
        data.test_neg_edge_index = neg[:, n_v:n_v + n_t]
        data.train_neg_edge_index = neg[:, n_v + n_t:]
    return {
        "train": {"edge": data.train_pos_edge_index.t(), "edge_neg": data.train_neg_edge_index.t()},
        "valid": {"edge": data.val_pos_edge_index.t(), "edge_neg": data.val_neg_edge_index.t()},
        "test": {"edge": data.test_pos_edge_index.t(), "edge_neg": data.test_neg_edge_index.t()},
    }


def split_edges(edge_index: torch.Tensor, val_ratio: float, test_ratio: float):
    """``src/generate_production_split.py:14-30``: one direction per undirected pair (self loops kept: ``<=``) is shuffled
    and cut into train / val / test; train and val come back symmetrised, test one-directional."""
    mask = edge_index[0] <= edge_index[1]
    perm = mask.nonzero(as_tuple=False).view(-1)
    perm = perm[torch.randperm(perm.size(0), device=perm.device)]
    num_val = int(val_ratio * perm.numel())
    num_test = int(test_ratio * perm.numel())
    num_train = perm.numel() - num_val - num_test
    train_edges = perm[:num_train]
    val_edges = perm[num_train:num_train + num_val]
    test_edges = perm[num_train + num_val:]
    train_ei = edge_index[:, train_edges]
    train_ei = torch.cat([train_ei, train_ei.flip([0])], dim=-1)
    val_ei = edge_index[:, val_edges]
    val_ei = torch.cat([val_ei, val_ei.flip([0])], dim=-1)
    return train_ei, val_ei, edge_index[:, test_edges]


def do_production_edge_split(dataset, data_name: str, test_ratio: float, val_node_ratio: float, val_ratio: float,
                             old_old_extra_ratio: float, split_seed: int = 234, verbose: bool = False):
    """``src/generate_production_split.py:32-95``: hide ``val_node_ratio`` of the nodes ("new" nodes), split old-old /
    old-new / new-new edges into training / inference / testing sets and draw one global set of undirected negatives.
    Returns ``(training_data, val_data, inference_data, data, test_edge_bundle, negative_samples)`` — the tuple
    ``train_teacher_gnn.py:348,366`` loads from ``../data/<ds>_production.pkl``."""
    random.seed(split_seed)
    torch.manual_seed(split_seed)
    assert len(dataset) == 1
    data = dataset[0]
    num_negatives = round(test_ratio * data.edge_index.size(1) / 2)
    negative_samples = negative_sampling(data.edge_index, data.num_nodes, num_negatives, force_undirected=True)
    # step 1: nodes to remove
    new_data = RandomNodeSplit(num_val=0.0, num_test=val_node_ratio)(data)
    # step 2: old-old edges -> training, extra inference edges, testing
    rows, cols = new_data.edge_index
    old_old = new_data.train_mask[rows] & new_data.train_mask[cols]
    old_old_train, old_old_val, old_old_test = split_edges(new_data.edge_index[:, old_old], old_old_extra_ratio, test_ratio)
    # step 3: old-new edges -> inference, testing
    old_new = (new_data.train_mask[rows] & new_data.test_mask[cols]) | (new_data.test_mask[rows] & new_data.train_mask[cols])
    old_new_train, _, old_new_test = split_edges(new_data.edge_index[:, old_new], 0.0, test_ratio)
    # step 4: new-new edges -> inference, testing
    new_new = new_data.test_mask[rows] & new_data.test_mask[cols]
    new_new_train, _, new_new_test = split_edges(new_data.edge_index[:, new_new], 0.0, test_ratio)
    # step 5: testing edges
    test_edge_index = torch.cat([old_old_test, old_new_test, new_new_test], dim=-1)
    test_edge_bundle = (old_old_test, old_new_test, new_new_test, test_edge_index)
    # step 6: the training graph lives on the old nodes only (relabelled)
    training_only_ei = subgraph(new_data.train_mask, old_old_train, relabel_nodes=True)[0]
    training_only_x = new_data.x[new_data.train_mask]
    # step 7: training / validation link split
    given_data = Data(training_only_x, training_only_ei)
    training_data, _, val_data = RandomLinkSplit(0.0, val_ratio, is_undirected=True)(given_data)
    # step 8: inference graph
    inference_edge_index = torch.cat([old_old_train, old_old_val, old_new_train, new_new_train], dim=-1)
    inference_data = Data(new_data.x, inference_edge_index)
    if verbose:
        print(f"Datasets Infomation:\\t\\nName:\\t{data_name}\\n#Old Nodes:\\t{training_only_x.size(0)}\\n"
              f"#New Nodes:\\t{new_data.x.size(0) - training_only_x.size(0)}\\n#Old-Old testing edges:\\t{old_old_test.size(1)}\\n"
              f"#Old-New testing edges:\\t{old_new_test.size(1)}\\n#New-New testing edges:\\t{new_new_test.size(1)}\\n")
    return training_data, val_data, inference_data, data, test_edge_bundle, negative_samples
