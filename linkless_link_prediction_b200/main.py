"""Student (LLP relational distillation) training — the reference's ``src/main.py`` surface over the B200 kernels.

* ``cosine_loss`` / ``kl_loss``  — src/main.py:24-31
* ``neighbor_samplers``          — :33-50
* ``train_minibatch``            — :52-144
* ``train``                      — :147-236
* ``main``                       — :238-513 (flags :240-269); unlike the reference, importing this module does
  NOT run ``main()`` (the reference calls it at import, :515).
"""
from __future__ import annotations

import argparse

import torch
import torch.nn.functional as F

from . import ops
from .loader import shuffled_batches
from .logger import Logger, ProductionLogger
from .models import MLP, LinkPredictor
from .optim import FusedAdam
from .shims import Evaluator, negative_sampling, random_walk, seed_everything
from . import shims
from . import train_teacher_gnn as _teacher
from .train_teacher_gnn import (CapturedStep, _captured_step_for, _dist, _shard, finish_distributed,
                                init_device, load_production, load_transductive, optimizer_tail, test_production,
                                test_transductive)


def cosine_loss(s, t):
    """KD_RM baseline (weight 0 by default; left to torch — out of the north-star scope)."""
    return 1 - F.cosine_similarity(s.float(), t.detach().float(), dim=-1).mean()


def kl_loss(s, t, T):
    """LLP_D: fused softmax/KL kernel (value == F.kl_div(log_softmax(s/T), softmax(t/T), 'sum') * T^2 / B)."""
    return ops.kl_loss(s, t, T)


def neighbor_samplers(row, col, sample, x, step, ps_method, ns_rate, hops):
    """Context nodes of each anchor: ``step`` uniform walks of ``hops`` ('nb') or one walk of ``step*hops`` ('rw')
    plus ``step*hops*ns_rate`` uniformly random nodes drawn with the CPU generator (main.py:47)."""
    batch = sample
    n = x.size(0)   # every node id is < n: spares torch_cluster's three blocking max() reductions per walk
    if ps_method == 'rw':
        pos_batch = random_walk(row, col, batch, walk_length=step * hops, coalesced=False, num_nodes=n)
    elif ps_method == 'nb':
        pos_batch = None
        for _ in range(step):
            w = random_walk(row, col, batch, walk_length=hops, coalesced=False, num_nodes=n)
            pos_batch = w if pos_batch is None else torch.cat((pos_batch, w[:, 1:]), 1)
    neg_batch = torch.randint(0, x.size(0), (batch.numel(), step * hops * ns_rate), dtype=torch.long)
    if batch.is_cuda:   # pinned + asynchronous: a pageable copy would block the host until the stream's earlier work (the
        neg_batch = neg_batch.pin_memory().to(batch.device, non_blocking=True)   # whole previous step) has finished
    return pos_batch.to(batch.device), neg_batch.to(batch.device)


def _score_kd_and_edges(predictor, h, anchor, ctx, src, dst):
    """Student scores of the (anchor, context) pairs AND of the train edges in ONE scorer call over the concatenated pair
    list: ``h`` then receives a single gradient from the scorer's backward (two calls make autograd add two full-size
    ``[rows, H]`` gradients) and the predictor's layers run once over all pairs.  Returns ``(s_r [B_n, K] or None,
    out [2B] or None)`` — the same values ``predictor(h[a].repeat(K), h[ctx])`` (main.py:184-186) and
    ``predictor(h[src], h[dst])`` (:212-214) give row for row."""
    n_kd = 0 if anchor is None else anchor.numel()
    n_e = 0 if src is None else src.numel()
    if n_kd == 0 and n_e == 0:
        return None, None
    if n_kd == 0:
        return None, predictor.score(h, src.contiguous(), dst.contiguous()).reshape(-1)
    if n_e == 0:
        return predictor.score(h, anchor, ctx).reshape(anchor.shape), None
    u = torch.cat((anchor.reshape(-1), src))
    v = torch.cat((ctx.reshape(-1), dst))
    s_r, out = torch.split(predictor.score(h, u, v).reshape(-1), (n_kd, n_e))
    return s_r.reshape(anchor.shape), out


def _teacher_kd_scores(teacher_predictor, t_h, anchor, ctx):
    with torch.no_grad():
        # the teacher predictor is never put in eval() by the reference: dropout stays active
        return teacher_predictor.score(t_h, anchor, ctx).reshape(anchor.shape)


def student_step(model, predictor, t_h, teacher_predictor, x, optimizer, args, edge, neg_edge, samples=None, node_perm=None,
                 kd_weight=1.0, edge_share=1.0):
    """One optimisation step of the full-batch student loop (main.py:169-230) on explicit device tensors: ``edge`` /
    ``neg_edge`` ``[2, B]`` (this rank's shard), ``samples`` ``[B_n, 1+K]`` = anchor | walk contexts | random contexts (or
    None when ``LLP_D == LLP_R == 0``), ``node_perm`` (only read when ``KD_RM`` != 0).  ``kd_weight`` / ``edge_share``
    are this rank's share of the global anchor / edge batch times W (1 on a single GPU).  A pure function of device
    tensors: ``StudentCapturedStep`` replays it as one CUDA graph.  Returns the (device) loss."""
    optimizer.zero_grad()
    ops.advance_rng(x.device)
    h = model(x)
    train_edges = torch.cat((edge, neg_edge), dim=-1)
    anchor = ctx = None
    if samples is not None and samples.size(0) > 0:
        K = samples.size(1) - 1
        anchor = samples[:, :1].expand(-1, K).contiguous()
        ctx = samples[:, 1:].contiguous()
    have_edges = train_edges.size(1) > 0
    # predictor(h[a].repeat(K), h[ctx]) (main.py:183-186) and predictor(h[src], h[dst]) (:212-214): one fused call
    s_r, out = _score_kd_and_edges(predictor, h, anchor, ctx, train_edges[0] if have_edges else None,
                                   train_edges[1] if have_edges else None)
    kd_total = None
    if samples is not None:
        if s_r is not None:
            t_r = _teacher_kd_scores(teacher_predictor, t_h, anchor, ctx)
            kd_total, _, _ = ops.kd_losses(s_r, t_r, 1, args.margin, args.LLP_D * kd_weight, args.LLP_R * kd_weight)
        else:   # empty anchor shard (fewer anchors than ranks): contributes nothing
            kd_total = torch.zeros((), dtype=torch.float32, device=h.device)
    if out is not None:
        label_loss = ops.bce_loss(out, edge.size(1))
        if edge_share != 1.0:
            label_loss = label_loss * edge_share
    else:   # empty shard: zero loss that still reaches backward / the gradient all-reduce through h
        label_loss = h.float().sum() * 0.0
    loss = args.True_label * label_loss
    if args.KD_RM:  # baselines, weight 0 by default; the reference evaluates them regardless
        # a mean over the node batch: every rank evaluates the whole (replicated) batch, so no re-weighting is needed
        loss = loss + args.KD_RM * cosine_loss(h[node_perm], t_h[node_perm])
    if args.KD_LM and out is not None:
        with torch.no_grad():
            t_out = teacher_predictor.score(t_h, train_edges[0].contiguous(), train_edges[1].contiguous()).reshape(-1)
        loss = loss + args.KD_LM * F.mse_loss(out, t_out) * edge_share   # a mean over this rank's edge shard
    if kd_total is not None:
        loss = loss + kd_total
    loss.backward()
    optimizer_tail(model, predictor, optimizer)
    return loss.detach()


class StudentCapturedStep(CapturedStep):
    """``student_step`` for fixed batch shapes as one CUDA-graph replay: ``step(edge, neg_edge, samples[, node_perm])``.
    Sampling (walks, random contexts, negative edges — host / torch RNG in the reference's order) stays outside the graph;
    everything from the encoder forward to Adam is inside."""

    def __init__(self, model, predictor, t_h, teacher_predictor, data, optimizer, args, kd_weight=1.0, edge_share=1.0,
                 eager_steps=2, profile=False):
        def fn(edge, neg_edge, samples=None, node_perm=None):
            return student_step(model, predictor, t_h, teacher_predictor, data.x, optimizer, args, edge, neg_edge, samples,
                                node_perm, kd_weight, edge_share)

        super().__init__(fn, (model, predictor), eager_steps, profile, lambda: [ops.to_compute(data.x, cache=True), t_h])


def train(model, predictor, t_h, teacher_predictor, data, split_edge, optimizer, args, device):
    if args.transductive == "transductive":
        pos_train_edge = split_edge['train']['edge'].to(data.x.device)
        row, col = data.adj_t
    else:
        pos_train_edge = data.edge_index.t()
        row, col = data.edge_index
    edge_index = torch.stack([col, row], dim=0)
    dev = data.x.device
    rank, world = _dist()

    model.train()
    predictor.train()

    total_loss = torch.zeros((), dtype=torch.float32, device=dev)
    total_examples = 0
    node_loader = shuffled_batches(data.x.size(0), args.node_batch_size * world)
    for link_perm in shuffled_batches(pos_train_edge.size(0), args.link_batch_size * world):
        node_perm = next(node_loader).to(dev)
        edge = pos_train_edge[link_perm.to(dev)].t()

        samples, kd_weight = None, 1.0
        if args.LLP_R or args.LLP_D:
            pos_sample, neg_sample = neighbor_samplers(row, col, node_perm, data.x, args.rw_step, args.ps_method,
                                                       args.ns_rate, args.hops)
            samples = torch.cat((pos_sample, neg_sample), 1)
            a_lo, a_hi = _shard(samples.size(0), rank, world)
            if world > 1:  # both terms are means over anchors: weight the shard by its share
                kd_weight = (a_hi - a_lo) * world / float(samples.size(0))
            samples = samples[a_lo:a_hi].contiguous()

        if args.datasets != "collab":
            neg_edge = negative_sampling(edge_index, num_nodes=data.x.size(0), num_neg_samples=link_perm.size(0),
                                         method='dense')
        else:
            neg_edge = torch.randint(0, data.x.size()[0], [edge.size(0), edge.size(1)], dtype=torch.long, device=dev)

        n_global = edge.size(1)
        lo, hi = _shard(n_global, rank, world)
        nlo, nhi = _shard(neg_edge.size(1), rank, world)
        edge, neg_edge = edge[:, lo:hi].contiguous(), neg_edge[:, nlo:nhi].contiguous()
        # this rank's share of the global batch times W: rank means average to the global mean
        edge_share = (edge.size(1) + neg_edge.size(1)) * world / float(2 * n_global) if world > 1 else 1.0

        tensors = [edge, neg_edge] + ([samples] if samples is not None else []) + \
                  ([node_perm] if (args.KD_RM and samples is not None) else [])
        step = None
        capturable = edge.size(1) + neg_edge.size(1) > 0 and (samples is None or samples.size(0) > 0) and \
            (samples is not None or not args.KD_RM)
        if _teacher.USE_CUDA_GRAPH and isinstance(optimizer, FusedAdam) and capturable:
            key = ("student", id(model), id(predictor), id(data), id(t_h), tuple(tuple(t.shape) for t in tensors),
                   float(kd_weight), float(edge_share), ops.compute_dtype(),
                   tuple(float(getattr(args, k)) for k in ("True_label", "KD_RM", "KD_LM", "LLP_D", "LLP_R", "margin")))
            step = _captured_step_for(optimizer, key, lambda: StudentCapturedStep(
                model, predictor, t_h, teacher_predictor, data, optimizer, args, kd_weight, edge_share))
        if step is not None:
            loss = step(*tensors)
        else:
            loss = student_step(model, predictor, t_h, teacher_predictor, data.x, optimizer, args, edge, neg_edge, samples,
                                node_perm, kd_weight, edge_share)

        total_loss += loss * n_global
        total_examples += n_global

    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(total_loss)
        total_loss /= world
    return total_loss.item() / total_examples


ENCODE_EVERY_NODE_ONCE = True   # feature-minibatch loop: see encode_every_node_once


def encode_every_node_once(model, num_nodes: int, rows_touched: int) -> bool:
    """Feature-minibatch step (main.py:93-101): may the encoder run once over all ``num_nodes`` rows instead of over the
    ``rows_touched`` gathered rows (duplicates included)?  Only when that is less work AND gives the same embeddings row for
    row: the student MLP is row-wise and deterministic when it has no norm layer and its dropout is inactive (every
    ``--minibatch`` configuration of the reference's scripts sets ``--dropout=0.0``, scripts/LLP_transductive.sh:5-8)."""
    if not ENCODE_EVERY_NODE_ONCE or rows_touched < num_nodes:
        return False
    if getattr(model, "norm_type", "none") != "none":
        return False
    drop = getattr(model, "dropout", None)
    p = float(getattr(drop, "p", drop if isinstance(drop, (int, float)) else 1.0))
    return (not model.training) or p == 0.0


def student_minibatch_step(model, predictor, t_h, teacher_predictor, x, optimizer, args, edge, neg_edge, samples,
                           kd_weight=1.0, edge_share=1.0):
    """One step of the feature-minibatch student loop (main.py:75-139) on explicit device tensors: only the rows the step
    touches — ``[samples.flatten(), src, dst]`` (:93-101) — go through the encoder; the scorers index that compact
    embedding matrix by position (or, when the touched rows with their duplicates outnumber the nodes and the encoder is
    deterministic, every node goes through it once and the scorers index by node id: ``encode_every_node_once``).  ``x`` stays resident in HBM (the reference keeps it on the host and copies the rows
    every step, :95-96), so the row gather is a device gather.  Capturable (``StudentCapturedStep``)."""
    optimizer.zero_grad()
    dev = x.device
    ops.advance_rng(dev)
    train_edges = torch.cat((edge, neg_edge), dim=-1)
    src, dst = train_edges[0], train_edges[1]
    n_s = samples.numel()
    K = samples.size(1) - 1
    anchor = ctx = src_pos = dst_pos = None
    if encode_every_node_once(model, x.size(0), n_s + 2 * src.numel()):
        # The step touches at least as many rows (duplicates included: every anchor K + 1 times, hub contexts over and
        # over) as the graph has nodes, and the encoder is deterministic and row-wise (no dropout, no norm): encoding
        # every node ONCE and indexing the scorers by node id gives, row for row, the embeddings the reference's
        # `model(x[this_target])` (main.py:95-101) produces, and the same parameter gradients (the duplicates' upstream
        # gradients are summed before the weight gradient instead of inside it).  collab: 747k rows -> 236k.
        h = model(x)
        if samples.size(0) > 0:
            anchor = samples[:, :1].expand(-1, K).contiguous()
            ctx = samples[:, 1:].contiguous()
        if src.numel() > 0:
            src_pos, dst_pos = src.contiguous(), dst.contiguous()
    else:
        this_target = torch.cat((samples.reshape(-1), src, dst), 0)
        h = model(x[this_target])  # rows of the touched nodes only
        if samples.size(0) > 0:
            local = torch.arange(n_s, device=dev).reshape(samples.shape)   # positions inside h of every sample
            anchor = local[:, :1].expand(-1, K).contiguous()
            ctx = local[:, 1:].contiguous()
        if src.numel() > 0:
            src_pos = torch.arange(n_s, n_s + src.numel(), device=dev)
            dst_pos = src_pos + src.numel()
    s_r, out = _score_kd_and_edges(predictor, h, anchor, ctx, src_pos, dst_pos)   # one scorer call: one gradient into h
    if s_r is not None:
        t_r = _teacher_kd_scores(teacher_predictor, t_h, samples[:, :1].expand(-1, K).contiguous(), samples[:, 1:].contiguous())
        kd_total, _, _ = ops.kd_losses(s_r, t_r, 1, args.margin, args.LLP_D * kd_weight, args.LLP_R * kd_weight)
    else:
        kd_total = h.float().sum() * 0.0
    if out is not None:
        label_loss = ops.bce_loss(out, edge.size(1))
        if edge_share != 1.0:
            label_loss = label_loss * edge_share
    else:
        label_loss = h.float().sum() * 0.0
    loss = args.True_label * label_loss + kd_total   # no KD_RM / KD_LM terms in this variant (main.py:129-130)
    loss.backward()
    optimizer_tail(model, predictor, optimizer)
    return loss.detach()


class StudentMinibatchCapturedStep(CapturedStep):
    """``student_minibatch_step`` for fixed batch shapes as one CUDA-graph replay: ``step(edge, neg_edge, samples)``."""

    def __init__(self, model, predictor, t_h, teacher_predictor, x, optimizer, args, kd_weight=1.0, edge_share=1.0,
                 eager_steps=2, profile=False):
        def fn(edge, neg_edge, samples):
            return student_minibatch_step(model, predictor, t_h, teacher_predictor, x, optimizer, args, edge, neg_edge,
                                          samples, kd_weight, edge_share)

        super().__init__(fn, (model, predictor), eager_steps, profile, lambda: [x, t_h])


def train_minibatch(model, predictor, t_h, teacher_predictor, data, split_edge, optimizer, args, device):
    """Feature-minibatch variant (main.py:52-144): only the rows a step touches are encoded.  On a 180 GB B200 the
    features stay resident in HBM (the reference keeps them on the host and copies rows every step, :95-96;
    SURVEY.md N4), so the per-step gather is a device gather.  Under ``torchrun`` the anchors and the edge batch of a
    (global) step are sharded across the ranks exactly as in ``train``."""
    if args.transductive == "transductive":
        pos_train_edge = split_edge['train']['edge'].to(device)
        row, col = data.adj_t
    else:
        pos_train_edge = data.edge_index.t().to(device)
        row, col = data.edge_index
    row, col = row.to(device), col.to(device)
    edge_index = torch.stack([col, row], dim=0)
    x = data.x.to(device)
    t_h = t_h.to(device)
    if not (args.LLP_D or args.LLP_R):
        raise NameError("name 'loss' is not defined")  # what the reference raises for this flag combination (:129-132)
    rank, world = _dist()

    model.train()
    predictor.train()
    total_loss = torch.zeros((), dtype=torch.float32, device=device)
    total_examples = 0
    node_loader = shuffled_batches(x.size(0), args.node_batch_size * world)
    for link_perm in shuffled_batches(pos_train_edge.size(0), args.link_batch_size * world):
        node_perm = next(node_loader).to(device)
        edge = pos_train_edge[link_perm.to(device)].t()
        if args.datasets != "collab":
            neg_edge = negative_sampling(edge_index, num_nodes=x.size(0), num_neg_samples=link_perm.size(0),
                                         method='dense')
        else:
            neg_edge = torch.randint(0, x.size()[0], [edge.size(0), edge.size(1)], dtype=torch.long).pin_memory().to(
                device, non_blocking=True)   # CPU generator as the reference (main.py:83-84); asynchronous copy

        pos_sample, neg_sample = neighbor_samplers(row, col, node_perm, x, args.rw_step, args.ps_method, args.ns_rate,
                                                   args.hops)
        samples = torch.cat((pos_sample, neg_sample), 1)
        n_global = edge.size(1)
        kd_weight = edge_share = 1.0
        if world > 1:
            a_lo, a_hi = _shard(samples.size(0), rank, world)
            kd_weight = (a_hi - a_lo) * world / float(samples.size(0))
            samples = samples[a_lo:a_hi]
            lo, hi = _shard(n_global, rank, world)
            nlo, nhi = _shard(neg_edge.size(1), rank, world)
            edge, neg_edge = edge[:, lo:hi], neg_edge[:, nlo:nhi]
            edge_share = (edge.size(1) + neg_edge.size(1)) * world / float(2 * n_global)
        edge, neg_edge, samples = edge.contiguous(), neg_edge.contiguous(), samples.contiguous()

        step = None
        if _teacher.USE_CUDA_GRAPH and isinstance(optimizer, FusedAdam) and samples.size(0) > 0 and edge.size(1) + neg_edge.size(1) > 0:
            key = ("student_mb", id(model), id(predictor), id(data), id(t_h), tuple(edge.shape), tuple(neg_edge.shape),
                   tuple(samples.shape), float(kd_weight), float(edge_share), ops.compute_dtype(),
                   tuple(float(getattr(args, k)) for k in ("True_label", "LLP_D", "LLP_R", "margin")))
            step = _captured_step_for(optimizer, key, lambda: StudentMinibatchCapturedStep(
                model, predictor, t_h, teacher_predictor, x, optimizer, args, kd_weight, edge_share))
        if step is not None:
            loss = step(edge, neg_edge, samples)
        else:
            loss = student_minibatch_step(model, predictor, t_h, teacher_predictor, x, optimizer, args, edge, neg_edge,
                                          samples, kd_weight, edge_share)
        total_loss += loss * n_global
        total_examples += n_global
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(total_loss)
        total_loss /= world
    return total_loss.item() / total_examples


def build_parser():
    parser = argparse.ArgumentParser(description='OGBL-DDI (GNN)')
    parser.add_argument('--device', type=int, default=0)
    parser.add_argument('--log_steps', type=int, default=1)
    parser.add_argument('--encoder', type=str, default='sage')
    parser.add_argument('--num_layers', type=int, default=2)
    parser.add_argument('--hidden_channels', type=int, default=256)
    parser.add_argument('--dropout', type=float, default=0.5)
    parser.add_argument('--link_batch_size', type=int, default=64 * 1024)
    parser.add_argument('--node_batch_size', type=int, default=64 * 1024)
    parser.add_argument('--lr', type=float, default=0.005)
    parser.add_argument('--epochs', type=int, default=20000)
    parser.add_argument('--eval_steps', type=int, default=5)
    parser.add_argument('--runs', type=int, default=10)
    parser.add_argument('--dataset_dir', type=str, default='../data')
    parser.add_argument('--datasets', type=str, default='collab')
    parser.add_argument('--predictor', type=str, default='mlp', choices=['inner', 'mlp'])
    parser.add_argument('--patience', type=int, default=100, help='number of patience steps for early stopping')
    parser.add_argument('--metric', type=str, default='Hits@20', choices=['auc', 'hits@20', 'hits@50'],
                        help='main evaluation metric')
    parser.add_argument('--use_valedges_as_input', action='store_true')
    parser.add_argument('--True_label', default=0.1, type=float, help="true_label loss")
    parser.add_argument('--KD_RM', default=0, type=float, help="Representation-based matching KD")
    parser.add_argument('--KD_LM', default=0, type=float, help="logit-based matching KD")
    parser.add_argument('--LLP_D', default=1, type=float, help="distribution-based matching kd")
    parser.add_argument('--LLP_R', default=1, type=float, help="rank-based matching kd")
    parser.add_argument('--margin', default=0.1, type=float, help="margin for rank-based kd")
    parser.add_argument('--rw_step', type=int, default=3, help="nearby nodes sampled times")
    parser.add_argument('--ns_rate', type=int, default=1, help="randomly sampled rate over # nearby nodes")
    parser.add_argument('--hops', type=int, default=2, help="random_walk step for each sampling time")
    parser.add_argument('--ps_method', type=str, default='nb', help="positive sampling is rw or nb")
    parser.add_argument('--transductive', type=str, default='transductive', choices=['transductive', 'production'])
    parser.add_argument('--minibatch', action='store_true')
    parser.add_argument('--precision', type=str, default='bf16', choices=['bf16', 'fp32'])
    parser.add_argument('--synthetic_scale', type=float, default=1.0)
    return parser


def main(argv=None):
    import os
    from .data import synthetic_dataset
    args = build_parser().parse_args(argv)
    print(args)
    ops.set_compute_dtype(args.precision)
    os.makedirs("../results", exist_ok=True)
    Logger_file = "../results/" + args.datasets + "_KD_" + args.transductive + ".txt"
    if int(os.environ.get("RANK", "0")) == 0:
        with open(Logger_file, "a") as file:
            file.write(str(args) + "\n")
            if args.KD_RM != 0:
                file.write("Logit-matching\n")
            elif args.KD_LM != 0:
                file.write("Representation-matching\n")
            elif args.LLP_D != 0 or args.LLP_R != 0:
                file.write("LLP (Relational Distillation)\n")
    device, rank, world = init_device(args)

    production = args.transductive != "transductive"
    if not production:
        # the SAME loader as the teacher driver: a split cached next to the data is the one the teacher's saved features
        # and predictor were trained on, so the student's valid/test edges can never be teacher training edges
        data, split_edge = load_transductive(args, device)
        input_size = data.x.size(1)
        args.metric = 'Hits@50' if args.datasets == "collab" else 'Hits@20'
        args.node_batch_size = int(data.x.size()[0] / (split_edge['train']['edge'].size()[0] / args.link_batch_size))
    else:   # main.py:337-348
        training_data, val_data, inference_data, test_edge_bundle, negative_samples = load_production(args, device)
        input_size = training_data.x.size(1)
        args.metric = 'Hits@20'
        args.node_batch_size = int(training_data.x.size()[0] / (training_data.edge_index.size(1) / args.link_batch_size))

    model = MLP(args.num_layers, input_size, args.hidden_channels, args.hidden_channels, args.dropout).to(device)
    predictor = LinkPredictor(args.predictor, args.hidden_channels, args.hidden_channels, 1, args.num_layers,
                              args.dropout).to(device)
    tag = args.datasets + "-" + args.encoder + "_" + args.transductive + ".pkl"
    pretrained_model = torch.load("../saved-models/" + tag, map_location=device)
    teacher_predictor = LinkPredictor(args.predictor, 256, 256, 1, 2, args.dropout)
    teacher_predictor.load_state_dict(pretrained_model['predictor'], strict=True)
    teacher_predictor.to(device)
    t_h = torch.load("../saved-features/" + tag, map_location=device)['features']
    for para in teacher_predictor.parameters():
        para.requires_grad = False

    evaluator = Evaluator(name='ogbl-ddi')
    keys = ['Hits@10', 'Hits@50', 'Hits@100', 'AUC'] if (args.datasets == "collab" and not production) else \
        ['Hits@10', 'Hits@20', 'Hits@30', 'Hits@50', 'AUC']
    loggers = {k: (ProductionLogger if production else Logger)(args.runs, args) for k in keys}

    for run in range(args.runs):
        seed_everything(run + 1)
        model.reset_parameters()
        predictor.reset_parameters()
        optimizer = FusedAdam(list(model.parameters()) + list(predictor.parameters()), lr=args.lr)
        cnt_wait, best_val = 0, 0.0
        for epoch in range(1, 1 + args.epochs):
            if not production:
                step_fn = train_minibatch if args.minibatch else train
                loss = step_fn(model, predictor, t_h, teacher_predictor, data, split_edge, optimizer, args, device)
                results, h = test_transductive(model, predictor, data, split_edge, evaluator, args.link_batch_size, 'mlp',
                                               args.datasets, args)
            else:   # main.py:419-423
                loss = train(model, predictor, t_h, teacher_predictor, training_data, None, optimizer, args, device)
                results, h = test_production(model, predictor, val_data, inference_data, test_edge_bundle, negative_samples,
                                             evaluator, args.link_batch_size, 'mlp', args.datasets)
            if results[args.metric][0] >= best_val:
                best_val, cnt_wait = results[args.metric][0], 0
            else:
                cnt_wait += 1
            for key, result in results.items():
                loggers[key].add_result(run, result)
            if epoch % args.log_steps == 0:
                for key, result in results.items():
                    print(key)
                    if not production:
                        valid_hits, test_hits = result
                        print(f'Run: {run + 1:02d}, Epoch: {epoch:02d}, Loss: {loss:.4f}, '
                              f'Valid: {100 * valid_hits:.2f}%, Test: {100 * test_hits:.2f}%')
                    else:
                        valid_hits, test_hits, old_old, old_new, new_new = result
                        print(f'Run: {run + 1:02d}, Epoch: {epoch:02d}, Loss: {loss:.4f}, valid: {100 * valid_hits:.2f}%, '
                              f'test: {100 * test_hits:.2f}%, old_old: {100 * old_old:.2f}%, old_new: {100 * old_new:.2f}%, '
                              f'new_new: {100 * new_new:.2f}%')
                print('---')
            if cnt_wait >= args.patience:
                break
        for key in loggers.keys():
            print(key)
            loggers[key].print_statistics(run)

    if world > 1:
        optimizer = None
        finish_distributed()
    if rank != 0:
        return
    with open(Logger_file, "a") as file:
        file.write('All runs:\n')
        for key in loggers.keys():
            print(key)
            loggers[key].print_statistics()
            file.write(f'{key}:\n')
            best_results = []
            for r in loggers[key].results:
                r = 100 * torch.tensor(r)
                best = r[:, 0].argmax()
                best_results.append(tuple(r[best, j].item() for j in range(r.size(1))) if production else
                                    (r[:, 0].max().item(), r[best, 1].item()))
            best_result = torch.tensor(best_results)
            if not production:
                r = best_result[:, 1]
                file.write(f'Test: {r.mean():.4f} ± {r.std():.4f}\n')
            else:   # main.py:488-511
                names = ('  Final val', '   Final Test', '   Final old_old', '   Final old_new', '   Final new_new')
                file.write(''.join(f'{nm}: {best_result[:, j].mean():.2f} ± {best_result[:, j].std():.2f}'
                                   for j, nm in enumerate(names)) + '\n')


if __name__ == "__main__":
    main()
