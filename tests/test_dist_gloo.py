"""Multi-rank host logic on CPU: world_size-2 ``gloo`` processes (the CUDA kernels themselves need a GPU; what is
covered here is everything around them that changes with the number of ranks — shard arithmetic, the loss weighting
that makes the averaged gradient equal the global-batch gradient, and the Hits@K candidate exchange)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from linkless_link_prediction_b200 import shims
from linkless_link_prediction_b200.train_teacher_gnn import _shard
from oracle import llp_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_topk(x, k):
    x = x.float()
    out = torch.full((k,), float("-inf"))
    v = torch.topk(x, min(k, x.numel())).values if x.numel() else x
    out[: v.numel()] = v
    return out


def _cpu_count(pos, thr):
    return (pos.float().unsqueeze(0) > thr.unsqueeze(1)).sum(1).to(torch.int64)


def _cpu_pairs(pos, neg):
    return torch.tensor(O.auc_pairs(pos, neg), dtype=torch.int64)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        # ---- Hits@K: sharded == unsharded, including ties and fewer negatives than K ----
        for n_pos, n_neg, Ks in ((1001, 5003, [10, 50, 100]), (37, 60, [10, 50, 100]), (5, 1, [1, 3])):
            pos = (torch.rand(n_pos, generator=g) * 50).round() / 50
            neg = (torch.rand(n_neg, generator=g) * 50).round() / 50
            lo, hi = _shard(n_pos, rank, world)
            nlo, nhi = _shard(n_neg, rank, world)
            counts, n = shims.hits_counts(pos[lo:hi], neg[nlo:nhi], Ks, group=dist.group.WORLD, topk_fn=_cpu_topk,
                                          count_fn=_cpu_count)
            assert counts.tolist() == O.hits_counts(pos, neg, Ks), (counts.tolist(), O.hits_counts(pos, neg, Ks))
            assert int(n) == n_pos
        # ---- ROC-AUC: sharded pair counts == unsharded (ragged shards, ties, an empty shard) ----
        for n_pos, n_neg in ((1001, 5003), (37, 60), (5, 1)):
            pos = (torch.rand(n_pos, generator=g) * 50).round() / 50
            neg = (torch.rand(n_neg, generator=g) * 50).round() / 50
            lo, hi = _shard(n_pos, rank, world)
            nlo, nhi = _shard(n_neg, rank, world)
            auc = shims.roc_auc_score_device(pos[lo:hi], neg[nlo:nhi], group=dist.group.WORLD, pairs_fn=_cpu_pairs)
            assert auc == O.roc_auc(pos, neg), (auc, O.roc_auc(pos, neg))
        # ---- node-partitioned encoder (ops.partition_messages): all-gather + local-row aggregation, forward
        # and transpose, reproduces the unpartitioned aggregation bit for bit (odd node count: padded last block) ----
        from linkless_link_prediction_b200.ops import partition_messages
        n_nodes, feat = 101, 8
        ei_p = O.synthetic_undirected_graph(n_nodes - 4, 500, seed=2)           # 4 isolated trailing nodes
        xg = torch.randn(n_nodes, feat, generator=torch.Generator().manual_seed(4))
        n_loc, lo, hi, (f_src, f_dst), (t_src, t_dst), inv_deg = partition_messages(ei_p, n_nodes, rank, world)
        assert n_loc == 51 and (lo, hi) == (rank * 51, rank * 51 + 51) and inv_deg.numel() == 102
        x_loc = torch.zeros(n_loc, feat); x_loc[: max(min(hi, n_nodes) - lo, 0)] = xg[lo:min(hi, n_nodes)]
        blocks = [torch.empty_like(x_loc) for _ in range(world)]
        dist.all_gather(blocks, x_loc)
        x_full = torch.cat(blocks)                                                # [N_padded, F]
        assert torch.equal(x_full[:n_nodes], xg)
        # forward rows of this rank: mean over the messages INTO the block
        rp, col, _ = O.csr_build(torch.stack([f_src, f_dst]), n_loc, "dst")
        agg_loc = O.spmm_csr(rp, col, x_full, mean=True)
        rp_g, col_g, _ = O.csr_build(ei_p, n_nodes, "dst")
        agg_ref = O.spmm_csr(rp_g, col_g, xg, mean=True)
        assert torch.equal(agg_loc[: min(hi, n_nodes) - lo], agg_ref[lo:min(hi, n_nodes)])
        # transpose rows of this rank: sum over the messages OUT of the block, scaled by the destination's 1/deg
        rp_t, col_t, _ = O.csr_build(torch.stack([t_dst, t_src]), n_loc, "dst")
        gt_loc = O.spmm_csr(rp_t, col_t, x_full, mean=False, src_scale=inv_deg)
        deg_g = torch.zeros(n_nodes).index_add_(0, ei_p[1], torch.ones(ei_p.size(1)))
        rp_tg, col_tg, _ = O.csr_build(ei_p, n_nodes, "src")
        gt_ref = O.spmm_csr(rp_tg, col_tg, xg, mean=False, src_scale=1.0 / deg_g.clamp(min=1))
        assert torch.equal(gt_loc[: min(hi, n_nodes) - lo], gt_ref[lo:min(hi, n_nodes)])
        # every message is owned exactly once per direction
        cnt = torch.tensor([f_src.numel(), t_src.numel()]); dist.all_reduce(cnt)
        assert cnt.tolist() == [ei_p.size(1), ei_p.size(1)]
        # ---- peer-memory variant (ops.stage_columns): each referenced remote row is staged ONCE behind the own block
        # and the re-coded column indices aggregate from [own block | staged rows] exactly what the all-gathered matrix
        # gives (the GPU path pulls the staged rows over NVLink; here they are read out of the gathered matrix) ----
        from linkless_link_prediction_b200.ops import stage_columns
        for (col_g, scale, mean, ref_out) in ((col, None, True, agg_loc), (col_t, inv_deg, False, gt_loc)):
            ref, local, sc = stage_columns(col_g, lo, hi, scale)
            assert torch.equal(ref, torch.unique(ref)) and bool(((ref < lo) | (ref >= hi)).all())      # sorted, distinct, remote
            assert int(local.max()) < n_loc + ref.numel() and int(local.min()) >= 0
            mat = torch.cat([x_full[lo:hi], x_full[ref]])                                          # [own block | staged rows]
            assert torch.equal(mat[local.long()], x_full[col_g.long()])                            # same row behind every edge
            rp_s = rp if mean else rp_t
            out = O.spmm_csr(rp_s, local.long(), mat, mean=mean, src_scale=sc)
            assert torch.equal(out, ref_out)
        # ---- sparse return of the embedding gradient: owners add, in rank order, the rows each rank touched in their
        # block == dense reduce-scatter (sum) of the [N_padded, F] matrices ----
        gi = torch.Generator().manual_seed(20 + rank)
        ids = torch.randint(0, n_nodes, (37,), generator=gi)
        g_full = torch.zeros(n_loc * world, feat)
        g_full[torch.unique(ids)] = torch.randn(torch.unique(ids).numel(), feat, generator=gi)
        published = [None] * world
        dist.all_gather_object(published, (ids, g_full))
        mine = torch.zeros(n_loc, feat)
        for q in range(world):                                   # rank order, marked rows only
            q_ids, q_g = published[q]
            marked = torch.unique(q_ids[(q_ids >= lo) & (q_ids < hi)])
            mine[marked - lo] += q_g[marked]
        total = g_full.clone(); dist.all_reduce(total)
        assert torch.equal(mine, total[lo:hi])
        # ---- training step: W ranks on shards of a 2B batch == 1 rank on the whole batch ----
        torch.manual_seed(0)
        n, f, H, B = 120, 16, 16, 101  # odd batch -> ragged shards
        ei = O.synthetic_undirected_graph(n, 400, seed=1)
        x = torch.randn(n, f)
        model = O.SAGE("t", f, H, H, 2, 0.0)
        pred = O.LinkPredictor("mlp", H, H, 1, 2, 0.0)
        edge = ei[:, :B]
        neg = torch.randint(0, n, (2, B))
        params = list(model.parameters()) + list(pred.parameters())

        def grads(e, ng, weight):
            for p in params:
                p.grad = None
            h = model(x, ei)
            te = torch.cat((e, ng), -1)
            out = pred(h[te[0]], h[te[1]]).squeeze()
            label = torch.cat((torch.ones(e.size(1)), torch.zeros(ng.size(1))))
            (O.bce_loss(out, label) * weight).backward()
            return torch.cat([p.grad.reshape(-1) for p in params])

        full = grads(edge, neg, 1.0)
        lo, hi = _shard(B, rank, world)
        weight = ((hi - lo) * 2) * world / float(2 * B)  # the rule in train_teacher_gnn.train
        local = grads(edge[:, lo:hi], neg[:, lo:hi], weight)
        dist.all_reduce(local)
        local /= world  # FusedAdam folds this 1/W into its kernel (grad_scale)
        torch.testing.assert_close(local, full, rtol=1e-5, atol=1e-7)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
