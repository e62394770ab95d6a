"""Data-parallel parity on real GPUs (run with torchrun, W >= 2):
W ranks x batch B with one NCCL all-reduce per step  ==  1 rank x batch W*B   (SURVEY.md §8e), and
sharded Hits@K == unsharded Hits@K (bit-exact counts).  Prints one JSON line from rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import linkless_link_prediction_b200 as L  # noqa: E402
from linkless_link_prediction_b200 import ops, shims  # noqa: E402
from linkless_link_prediction_b200 import train_teacher_gnn as teacher  # noqa: E402
from linkless_link_prediction_b200.data import synthetic_dataset  # noqa: E402


def build(dev, seed):
    shims.seed_everything(seed)
    model = L.SAGE("cora", 1433, 64, 64, 2, 0.0).to(dev)
    pred = L.LinkPredictor("mlp", 64, 64, 1, 2, 0.0).to(dev)
    return model, pred


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    teacher.EVAL_SHARD_MIN_EDGES = 0   # exercise the sharded Hits@K / AUC exchange even on these small edge lists
    ops.set_compute_dtype(torch.float32)
    data, split = synthetic_dataset("cora", seed=0)
    data = data.to(dev)
    B = 1024
    # (a) W ranks, per-rank batch B
    model, pred = build(dev, 0)
    opt = L.FusedAdam(list(model.parameters()) + list(pred.parameters()), lr=0.01)
    shims.seed_everything(1)
    loss_dp = [teacher.train(model, pred, data, split, opt, B, "sage", "cora", "transductive") for _ in range(2)]
    flat_dp = opt.flat_param.clone()
    args = type("A", (), {"minibatch": False, "compute_auc": False})()
    res_dp, _ = teacher.test_transductive(model, pred, data, split, L.Evaluator(), B, "sage", "cora", args)
    # (b) the same job on one rank with batch W*B (collectives disabled)
    real = teacher._dist
    teacher._dist = lambda: (0, 1)
    try:
        model1, pred1 = build(dev, 0)
        opt1 = L.FusedAdam(list(model1.parameters()) + list(pred1.parameters()), lr=0.01, distributed=False)
        shims.seed_everything(1)
        loss_1 = [teacher.train(model1, pred1, data, split, opt1, B * world, "sage", "cora", "transductive") for _ in range(2)]
        res_1, _ = teacher.test_transductive(model1, pred1, data, split, L.Evaluator(), B, "sage", "cora", args)
        # sharded eval of the SAME model as (b): counts must be identical
        teacher._dist = real
        res_1_sharded, _ = teacher.test_transductive(model1, pred1, data, split, L.Evaluator(), B, "sage", "cora", args)
    finally:
        teacher._dist = real
    diff = (flat_dp - opt1.flat_param).abs().max().item()
    scale = opt1.flat_param.abs().max().item()
    ok = diff <= 2e-5 * scale + 1e-6 and all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(loss_dp, loss_1)) \
        and res_1 == res_1_sharded
    if rank == 0:
        print(json.dumps({"world": world, "loss_dp": loss_dp, "loss_1rank_bigbatch": loss_1, "max_param_diff": diff,
                          "param_scale": scale, "hits_unsharded": res_1, "hits_sharded": res_1_sharded,
                          "hits_dp_model": res_dp, "ok": bool(ok)}))
    # CUDA graphs that captured the gradient all-reduce are still alive: release them first, then a regular teardown
    for o in (opt, opt1):
        o.__dict__.pop("_llp_captured_steps", None)
    del opt, opt1
    teacher.finish_distributed()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
