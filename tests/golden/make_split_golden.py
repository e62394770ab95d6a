"""Golden fixtures for the split generators, produced by the REFERENCE'S OWN functions.

Run here (needs ``/root/reference``; the tests only read the committed ``split_golden.pt``):

    python tests/golden/make_split_golden.py

``src/utils.py:62-105`` (``do_edge_split``) and ``src/generate_production_split.py:14-95`` (``split_edges``,
``do_production_edge_split``) are imported unmodified; the torch_geometric symbols they call are stubbed with the
ORACLE's restatements (``oracle/llp_oracle.py``, section N3 — written independently of the product's ``splits.py`` /
``shims.py``; torch_geometric itself is not installed, so that half stays "parity unpinned").  The fixture pins the
reference-owned glue — which edges go where, in which order the RNG streams are consumed, the container layout — and the
product, which re-implements glue AND third-party pieces, has to reproduce it index for index (``tests/test_splits.py``).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import _Stub  # noqa: E402
from linkless_link_prediction_b200.data import features, undirected_graph  # noqa: E402  (synthetic inputs only)
from oracle import llp_oracle as O  # noqa: E402


def install_stubs():
    names = ["torch_geometric", "torch_geometric.utils", "torch_geometric.transforms", "torch_geometric.data",
             "torch_geometric.datasets", "ogb", "ogb.linkproppred"]
    mods = {n: _Stub(n) for n in names}
    for n, m in mods.items():
        sys.modules[n] = m
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(mods[parent], child, m)
    u, t, d = mods["torch_geometric.utils"], mods["torch_geometric.transforms"], mods["torch_geometric.data"]
    u.negative_sampling = O.negative_sampling
    u.add_self_loops = O.add_self_loops
    u.train_test_split_edges = O.train_test_split_edges
    u.subgraph = O.subgraph
    t.RandomLinkSplit = O.RandomLinkSplit
    t.RandomNodeSplit = O.RandomNodeSplit
    d.Data = O.GraphData
    d.Dataset = list


def graph(n, pairs, f, seed):
    ei = undirected_graph(n, pairs, seed, True)
    return O.GraphData(x=features(n, f, 0.1, seed), edge_index=ei)


def pack_data(d):
    return {k: v for k, v in d.__dict__.items() if torch.is_tensor(v)}


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import generate_production_split as ref_prod
    import utils as ref_utils
    out = {}
    for tag, fast in (("edge_split", False), ("edge_split_fast", True)):
        data = graph(300, 1500, 8, 1)
        out[tag] = {"args": dict(n=300, pairs=1500, f=8, seed=1, fast_split=fast, val_ratio=0.05, test_ratio=0.1),
                    "split_edge": ref_utils.do_edge_split([data], fast_split=fast, val_ratio=0.05, test_ratio=0.1)}
    # the copy of do_edge_split in generate_production_split.py takes a split_seed
    data = graph(300, 1500, 8, 1)
    out["edge_split_seed7"] = {"args": dict(n=300, pairs=1500, f=8, seed=1, fast_split=False, val_ratio=0.1, test_ratio=0.2, split_seed=7),
                               "split_edge": ref_prod.do_edge_split([data], False, 0.1, 0.2, 7)}
    torch.manual_seed(3)
    ei = graph(120, 500, 4, 2).edge_index
    out["split_edges"] = {"edge_index": ei, "seed": 3, "val_ratio": 0.1, "test_ratio": 0.2,
                          "out": ref_prod.split_edges(ei, 0.1, 0.2)}
    data = graph(400, 2500, 8, 4)
    training_data, val_data, inference_data, full, bundle, negs = ref_prod.do_production_edge_split(
        [data], "synthetic", 0.1, 0.1, 0.1, 0.1)
    out["production"] = {"args": dict(n=400, pairs=2500, f=8, seed=4, test_ratio=0.1, val_node_ratio=0.1, val_ratio=0.1,
                                      old_old_extra_ratio=0.1),
                         "training_data": pack_data(training_data), "val_data": pack_data(val_data),
                         "inference_data": pack_data(inference_data), "test_edge_bundle": tuple(bundle),
                         "negative_samples": negs}
    path = os.path.join(HERE, "split_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
