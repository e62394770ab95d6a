#!/usr/bin/env python
"""Host-side wall time of the pieces of one LLP student step (no device syncs inside the timed pieces): where does the
host time of `bench.py --workload {cora,physics}-student` go?  Development tool.
    python tools/student_host_profile.py physics-student"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "physics-student"
    sys.argv = [sys.argv[0], "--workload", wl, "--steps", "5", "--warmup", "3", "--no-cpu-baseline"]
    args = bench.parse()
    hz = bench.Harness(args)
    torch.cuda.set_device(hz.dev)
    data_cpu, split = bench.build_workload(args)
    w = bench.make_student(hz, "bf16", data_cpu, split)
    from linkless_link_prediction_b200 import main as student, shims
    for _ in range(6):
        w["resident"]()
    torch.cuda.synchronize()
    # re-create the closure's pieces by hand
    import types
    cell = {n: c.cell_contents for n, c in zip(w["resident"].__code__.co_freevars, w["resident"].__closure__)}
    sas = cell["sample_and_step"]
    inner = {n: c.cell_contents for n, c in zip(sas.__code__.co_freevars, sas.__closure__)}
    a, row, col, data, dev, n_nodes, edge_index, step, cfg = (inner[k] for k in ("a", "row", "col", "data", "dev", "n_nodes", "edge_index", "step", "cfg"))
    pos_dev, batch = cell["pos_dev"], cell["batch"]
    T = {}
    def tick(name, t0):
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    reps = 50
    for _ in range(reps):
        t0 = time.perf_counter(); perm = torch.randint(0, pos_dev.size(0), (batch,), device=dev); edge = pos_dev[perm].t().contiguous(); tick("edge batch", t0)
        t0 = time.perf_counter(); node_perm = torch.randperm(n_nodes, device=dev)[:a.node_batch_size]; tick("randperm", t0)
        t0 = time.perf_counter(); ps, ns = student.neighbor_samplers(row, col, node_perm, data.x, a.rw_step, a.ps_method, a.ns_rate, a.hops); tick("neighbor_samplers", t0)
        t0 = time.perf_counter(); samples = torch.cat((ps, ns), 1); tick("cat", t0)
        t0 = time.perf_counter()
        if cfg["ds"] != "collab":
            neg = shims.negative_sampling(edge_index, num_nodes=n_nodes, num_neg_samples=edge.size(1), method="dense")
        else:
            neg = torch.randint(0, n_nodes, edge.size(), dtype=torch.long, device=dev)
        tick("negative_sampling", t0)
        t0 = time.perf_counter(); step(edge, neg, samples); tick("graph step launch", t0)
    torch.cuda.synchronize()
    tot = sum(T.values())
    for k, v in T.items():
        print(f"{k:22s} {1e3 * v / reps:8.3f} ms")
    print(f"{'total host':22s} {1e3 * tot / reps:8.3f} ms per step ({wl}; anchors {a.node_batch_size}, batch {batch})")


if __name__ == "__main__":
    main()
