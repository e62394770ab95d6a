#!/usr/bin/env python
"""Headline benchmark of the LLP hot path (BASELINE.json: "edges/sec: SAGE+LinkPredictor train & Hits@K eval scoring").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload collab]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE training step of the ogbl-collab-shaped teacher (BASELINE.json configs[3]: 235,868 nodes,
2,358,104 messages, 128-d features, SAGE 128->256->256->256 + LinkPredictor(256,256,1,2), dropout 0.5, B = 65,536
positive edges + 65,536 negatives per rank): full-graph encoder forward+backward, fused edge scoring, BCE,
(all-reduce at N>1), clip + Adam.  ``value`` = positive edges consumed per second over all ranks with inputs resident
in HBM; ``e2e`` = the same through the public ``train_step`` with the step's edge batch coming from pinned host memory
and the loss read back every step.  One eval pass (encoder forward + 4 scoring sets + Hits@{10,50,100}) is timed
separately and reported under ``eval``.  The CPU arm (``--impl reference`` and the ``cpu_baseline`` object) times the
pure-torch CPU oracle of the same step on the host cores: the reference itself cannot run here (its torch_geometric /
torch_scatter / torch_cluster / ogb dependencies are absent).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, LAYERS, BATCH, DROPOUT, LR = 256, 3, 65536, 0.5, 0.005   # LAYERS is per workload: see layers_of()


def layers_of(workload):
    """collab: SAGE 128->256->256->256 (scripts/supervised_transductive.sh); the synthetic power-law config
    (BASELINE.json configs[4]) is 256->256->256."""
    return 3 if workload == "collab" else 2

# dram__bytes_read.sum + dram__bytes_write.sum per SpMM launch (mean of one step's five launches) from the committed
# `ncu --set full` capture of this same command (tools/ncu_summary.py); algorithmic bytes per launch are 1.2 GB, the rest
# are L2 hits.  Only meaningful for the collab workload the capture was taken on.
SPMM_TRAFFIC_SOURCE = "profiles/r01_spmm_ncu_full_summary.json"


def spmm_traffic(workload):
    path = os.path.join(ROOT, SPMM_TRAFFIC_SOURCE)
    if workload != "collab" or not os.path.exists(path):
        return None
    try:
        t = [l["traffic_bytes"] for l in json.load(open(path))["launches"] if "spmm_kernel" in l["kernel"] and "traffic_bytes" in l]
        return sum(t) / len(t) if t else None
    except (OSError, ValueError, KeyError):
        return None



def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="collab")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic graph (debugging only)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-baseline-seconds", type=float, default=25.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="weight gradients on the main stream (ops.OVERLAP_WGRAD = False)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a thread (one sample every ~25 ms, so
    even a 40 ms region is covered), falling back to an `nvidia-smi -lms` child process when NVML is unavailable."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.sm, self.max_mhz, self.reasons, self.reason_samples = [], None, set(), 0
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _reasons(self, nv, handle):
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
        except AttributeError:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
        for name, bit in self.BITS.items():
            if mask & bit:
                self.reasons.add(name)
        self.reason_samples += 1

    def _nvml_loop(self, nv, handle):
        # every NVML query contends with kernel launches (see the sleep below), so the throttle-reason query is issued
        # with the first sample and whenever the SM clock is below its maximum; a clock AT its maximum is not throttled
        first = True
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM))
                self.sm.append(mhz)
                if first or (self.max_mhz and mhz < 0.98 * self.max_mhz):
                    self._reasons(nv, handle)
                first = False
            except Exception:  # noqa: BLE001 - a failed sample is just a missing sample
                pass
            time.sleep(0.025)   # (NVML queries contend with kernel launches: polling every 2 / 10 / 15 ms slowed 2-rank steps by ~20 / 14 / 11 %)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it lists integers
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.index]) if len(ids) > self.index else self.index
            handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self._nvml = nv
            self._handle = handle
            self._thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self._thread.start()
            return
        except Exception:  # noqa: BLE001 - fall back to nvidia-smi
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            try:  # one more reason sample right at the end of the region (sticky reasons such as sw_power_cap)
                self._reasons(self._nvml, self._handle)
            except Exception:  # noqa: BLE001
                pass
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(sm), "reason_samples": self.reason_samples, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def build_workload(args, seed=0):
    from linkless_link_prediction_b200.data import synthetic_dataset
    data, split = synthetic_dataset(args.workload, seed=seed, scale=args.scale)
    return data, split


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference step on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_runner(data, split, batch, layers, updated=False):
    from oracle import llp_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    x, adj = data.x, data.adj_t
    model = O.SAGE("collab", x.size(1), HIDDEN, HIDDEN, layers, DROPOUT, O.SAGEConvUpdated if updated else O.SAGEConv)
    pred = O.LinkPredictor("mlp", HIDDEN, HIDDEN, 1, 2, DROPOUT)
    opt = torch.optim.Adam(list(model.parameters()) + list(pred.parameters()), lr=LR)
    pos = split["train"]["edge"]
    model.train(); pred.train()
    g = torch.Generator().manual_seed(1)

    def step():
        perm = torch.randint(0, pos.size(0), (batch,), generator=g)
        edge = pos[perm].t()
        neg = torch.randint(0, x.size(0), edge.size(), dtype=torch.long, generator=g)
        return O.teacher_step(model, pred, x, adj, edge, neg, opt)

    return step


def run_cpu(args, data, split, budget_s, max_steps, warmup=1):
    step = cpu_step_runner(data, split, BATCH, layers_of(args.workload), updated=args.workload == "coauthor-physics")
    for _ in range(warmup):
        step()
    t0, n = time.perf_counter(), 0
    while n < max_steps and (n == 0 or time.perf_counter() - t0 < budget_s):
        step()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": BATCH * n / dt, "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} full training steps (B={BATCH}) of the same {args.workload}-shaped workload after {warmup} warm-up, "
                      f"pure-torch CPU oracle (index_select + index_add_ mean, fp32 nn.Linear)", "steps": n,
            "ms_per_step": 1e3 * dt / n}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    which = "configs[3]" if args.workload == "collab" else ("configs[4], scaled by %g" % args.scale if args.workload == "powerlaw-10m" else "shape table in data.py")
    config = {"workload": f"{args.workload}-shaped teacher train step (BASELINE.json {which})", "nodes": None,
              "messages": None, "feat": None, "hidden": HIDDEN, "layers": layers_of(args.workload), "batch_pos_edges_per_gpu": BATCH,
              "dropout": DROPOUT, "parallelism": f"dp{world}", "l2": "per-step working set (>1 GB) exceeds the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        data, split = build_workload(args)
        config.update(nodes=data.x.size(0), messages=data.adj_t.size(1), feat=data.x.size(1))
        res = run_cpu(args, data, split, budget_s=150.0, max_steps=max(args.steps, 1), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": "train_pos_edges_per_sec", "value": res["value"], "unit": "edges/s",
                "n_gpus": args.gpus, "gpus_used": 0, "steps": res["steps"], "warmup": min(args.warmup, 1), "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "reference cannot run here (torch_geometric/torch_scatter/torch_cluster/ogb absent): CPU oracle port timed"}
        print(json.dumps(line))
        return

    import linkless_link_prediction_b200 as L
    from linkless_link_prediction_b200 import _native as N
    from linkless_link_prediction_b200 import ops, shims
    from linkless_link_prediction_b200 import train_teacher_gnn as teacher

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    N.require_gpu()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ops.set_compute_dtype(args.precision)
    if args.no_overlap:
        ops.OVERLAP_WGRAD = set()

    data_cpu, split = build_workload(args)
    config.update(nodes=data_cpu.x.size(0), messages=data_cpu.adj_t.size(1), feat=data_cpu.x.size(1))
    data = shims.Data(x=data_cpu.x, adj_t=data_cpu.adj_t).to(dev)
    shims.seed_everything(0)
    # the reference picks SAGEConv_updated for coauthor-physics (train_teacher_gnn.py:376-379), PyG SAGEConv otherwise
    conv = L.SAGEConv_updated if args.workload == "coauthor-physics" else L.SAGEConv
    model = L.SAGE(args.workload, data.x.size(1), HIDDEN, HIDDEN, layers_of(args.workload), DROPOUT, conv).to(dev)
    predictor = L.LinkPredictor("mlp", HIDDEN, HIDDEN, 1, 2, DROPOUT).to(dev)
    optimizer = L.FusedAdam(list(model.parameters()) + list(predictor.parameters()), lr=LR)
    model.train(); predictor.train()
    pos_dev = split["train"]["edge"].to(dev)
    n_nodes = data.x.size(0)
    shims.seed_everything(1234 + rank)  # every rank trains on its own shard of the global batch
    total = args.warmup + args.steps

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- (1) device-resident timing: value ------------------------------------------------------
    # The step is replayed as ONE CUDA graph (teacher.CapturedTrainStep): the first two warm-up steps run eagerly, the
    # third captures.  Event-record nodes around every SpMM launch of the graph give the dominant kernel's duration on
    # the stream it runs on.
    step = teacher.CapturedTrainStep(model, predictor, data, optimizer, eager_steps=min(2, max(args.warmup - 1, 1)),
                                     profile_spmm=True)

    def step_resident():
        perm = torch.randint(0, pos_dev.size(0), (BATCH,), device=dev)
        edge = pos_dev[perm].t()
        neg = torch.randint(0, n_nodes, edge.size(), dtype=torch.long, device=dev)  # collab branch, train_teacher_gnn.py:53
        return step(edge, neg)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0, replays0 = N.launch_count(), step.replays
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps  # CPU time to enqueue one step (no sync inside)
    e1.record()
    barrier()
    launches = (N.launch_count() - launches0) + (step.replays - replays0) * step.launches_per_replay
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = BATCH * world * args.steps / (ms / 1e3)
    peak, peak_src = peaks()
    # SpMM durations: the event nodes hold the LAST timed step's launches now; then `steps` more replays, read one by one
    prof = step.spmm_events
    spmm_bytes_step = sum(nb for _, _, nb in prof)
    last_step_ms = sum(a.elapsed_time(b) for a, b, _ in prof) if prof else 0.0
    spmm_ms = 0.0
    for _ in range(args.steps if prof else 0):
        step_resident()
        torch.cuda.synchronize()
        spmm_ms += sum(a.elapsed_time(b) for a, b, _ in prof)
    spmm_launches = len(prof) * args.steps
    spmm_bytes = spmm_bytes_step * args.steps
    achieved = spmm_bytes / (spmm_ms / 1e3) / 1e9 if spmm_ms > 0 else None
    achieved_last = spmm_bytes_step / (last_step_ms / 1e3) / 1e9 if last_step_ms > 0 else None
    spmm_share = (spmm_ms / args.steps) / (ms / args.steps) if ms and args.steps else None

    # ---- (2) end-to-end through the public step with host inputs: e2e ---------------------------
    host_batches = [split["train"]["edge"][torch.randint(0, pos_dev.size(0), (BATCH,))].t().contiguous().pin_memory()
                    for _ in range(min(total, 8))]
    h2d = host_batches[0].numel() * host_batches[0].element_size()

    def step_e2e(i):
        edge = host_batches[i % len(host_batches)].to(dev, non_blocking=True)
        neg = torch.randint(0, n_nodes, edge.size(), dtype=torch.long, device=dev)
        loss = step(edge, neg)
        return loss.item()  # 4-byte D2H + sync every step, like the reference's loss.item() (train_teacher_gnn.py:70)

    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    e0.record()
    last = None
    for i in range(args.steps):
        last = step_e2e(i)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = BATCH * world * args.steps / (ms_e2e / 1e3)

    # ---- (3) eval pass: encoder forward + scoring of valid/test pos/neg + Hits@K ----------------
    ev_args = type("A", (), {"minibatch": False, "compute_auc": False})()
    n_scored = sum(split[k][j].size(0) for k in ("valid", "test") for j in ("edge", "edge_neg"))
    teacher.test_transductive(model, predictor, data, split, L.Evaluator(), BATCH, "sage", args.workload, ev_args)
    barrier()
    e0.record()
    results, _ = teacher.test_transductive(model, predictor, data, split, L.Evaluator(), BATCH, "sage", args.workload, ev_args)
    e1.record()
    barrier()
    ms_eval = max_over_ranks(e0.elapsed_time(e1))
    model.train(); predictor.train()

    def finish():
        """Leave without NCCL teardown: a CUDA graph that captured the gradient all-reduce is still alive, and
        destroying the communicator under it can block forever.  Everything has been synchronised by now."""
        if world > 1:
            import torch.distributed as dist
            step.graph = None
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return

    line = {
        "metric": "train_pos_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "host_enqueue_ms_per_step": host_enqueue_ms,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": e2e, "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": last},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "spmm_kernel (+fix-up), SAGE mean aggregation fwd + transpose-bwd",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None,
                     "traffic": spmm_traffic(args.workload), "traffic_source": SPMM_TRAFFIC_SOURCE,
                     "launches_timed": spmm_launches, "share_of_step": spmm_share,
                     "algorithmic_bytes_per_launch": spmm_bytes_step / max(len(prof), 1),
                     "algorithmic_bytes_per_step": spmm_bytes_step,
                     "achieved_last_timed_step": achieved_last,
                     "timing": "event-record nodes around each SpMM launch inside the step's CUDA graph, on the capture "
                               "stream; `steps` replays read one by one after the timed region (+ the last timed step)"},
        "eval": {"scored_edges_per_sec": n_scored * 1e3 / ms_eval, "ms": ms_eval, "scored_edges": n_scored,
                 "hits": {k: v for k, v in results.items()}},
    }
    if world == 1 and not args.no_cpu_baseline:
        res = run_cpu(args, data_cpu, split, budget_s=args.cpu_baseline_seconds, max_steps=3, warmup=1)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
