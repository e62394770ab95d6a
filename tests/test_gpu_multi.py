"""Multi-GPU parity inside pytest (skipped when fewer than two GPUs are visible; `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu` runs them).  Each test launches a 2-rank `torch.distributed.run` job over NCCL:

* data parallelism: W ranks x batch B with one gradient all-reduce per step == 1 rank x batch W*B, and sharded
  Hits@K == unsharded Hits@K (tools/dp_parity.py);
* node-partitioned encoder (ops.PartitionedGraph): embeddings, losses, gradients, parameters, Hits@K and
  AUC against the replicated single-GPU encoder in both precision modes (tools/np_parity.py).
The world_size-2 gloo tests in test_dist_gloo.py cover the same host-side arithmetic on the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_two_ranks(script, *extra, port=29617, timeout=600):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script), *extra]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, (res.returncode, res.stdout[-2000:], res.stderr[-3000:])
    return json.loads(lines[-1])


def test_data_parallel_equals_single_rank_big_batch(cuda):
    out = _run_two_ranks("dp_parity.py", port=29617)
    assert out["ok"] and out["world"] == 2
    assert out["hits_unsharded"] == out["hits_sharded"]


def test_node_partitioned_encoder_equals_replicated(cuda):
    out = _run_two_ranks("np_parity.py", port=29618)
    assert out["ok"] and out["fp32"]["eval_identical"] and out["bf16"]["eval_identical"]
    assert out["bf16"]["embeddings_bit_identical"]


@pytest.mark.parametrize("flag", ["--peer", "--peer-load"])
def test_node_partitioned_encoder_peer_memory_spmm(cuda, flag):
    """The same parity job with ``PartitionedGraph(peer=...)``: buffers mapped through CUDA IPC, ``llp_peer_barrier``
    around the NVLink traffic; ``--peer`` pulls every referenced remote row once (``llp_peer_gather_rows``) and
    aggregates locally, ``--peer-load`` lets the SpMM kernel load remote rows per edge (``llp_spmm_peer``).  Both are
    bit-identical to the all-gather SpMM of the same partition."""
    out = _run_two_ranks("np_parity.py", flag, port=29619 if flag == "--peer" else 29620)
    assert out["ok"] and out["fp32"]["peer"] and out["fp32"]["peer_spmm_bit_identical"] and out["bf16"]["peer_spmm_bit_identical"]
    assert not out["fp32"]["barrier_timed_out"] and not out["bf16"]["barrier_timed_out"]
