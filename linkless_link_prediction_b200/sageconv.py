"""SAGE convolution layers with the reference's module surface.

``SAGEConv`` mirrors ``torch_geometric.nn.SAGEConv(in, out)`` as the reference uses it
(``train_teacher_gnn.py:381-383``: mean aggregation, root weight, bias on ``lin_l`` only) and
``SAGEConv_updated`` mirrors ``src/sageconv_updated.py:9-93`` (transform, then aggregate).
State-dict keys (``lin_l.weight``, ``lin_l.bias``, ``lin_r.weight``; weight ``[out, in]``) and the
parameter initialisation order are the reference's, so checkpoints interchange.
``forward(x, edge_index)`` takes the dense ``[2,E]`` LongTensor the drivers pass.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _SageBase(nn.Module):
    _fn = None

    def __init__(self, in_channels: int, out_channels: int, normalize: bool = False, root_weight: bool = True,
                 bias: bool = True, **kwargs):
        super().__init__()
        if not root_weight or not bias:
            raise NotImplementedError("only root_weight=True, bias=True (what the reference drivers use) is built")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalize, self.root_weight = normalize, root_weight
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()

    def forward(self, x, edge_index, size=None, *, _relu: bool = False, _dropout: float = 0.0, _in_gate: float = 0.0,
                _defer_gate: bool = False, _input_layer: bool = False):
        """``_relu`` / ``_dropout`` are set by ``models.SAGE`` to fuse its relu + dropout into this layer's epilogue;
        ``_in_gate`` / ``_defer_gate`` move the relu/dropout backward mask into the consumer's GEMM epilogue (ops.py);
        ``_input_layer`` says that ``x`` is the constant feature matrix (its aggregate is loop-invariant, ops.Graph.spmm_input)."""
        if isinstance(x, (tuple, list)):
            x = x[0]
        graph = edge_index if isinstance(edge_index, ops.Graph) else ops.graph_of(edge_index, x.size(0))
        x = ops.to_compute(x, cache=True)
        p = float(_dropout) if self.training else 0.0
        seed, offset = ops._dropout_seed() if p > 0 else (0, 0)
        if self.normalize:
            # sageconv_updated.py:78-79 (and PyG SAGEConv): L2-normalise the layer output; relu / dropout asked for by the
            # caller then run after the normalisation through torch (never enabled by the reference drivers)
            out = type(self)._fn.apply(x, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, graph, False, 0.0, 0, 0,
                                       float(_in_gate), False, bool(_input_layer))
            out = torch.nn.functional.normalize(out.float(), p=2.0, dim=-1)
            if _relu:
                out = torch.relu(out)
            if _dropout > 0:
                out = torch.nn.functional.dropout(out, p=float(_dropout), training=self.training)
            return out.to(ops.compute_dtype()).contiguous()
        return type(self)._fn.apply(x, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, graph, bool(_relu), p, seed,
                                    offset, float(_in_gate), bool(_defer_gate), bool(_input_layer))

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels})"


class SAGEConv(_SageBase):
    """``lin_l(mean_{s->d} x[s]) + lin_r(x)`` — aggregate, then transform."""
    _fn = ops.SageConvFn


class SAGEConv_updated(_SageBase):
    """``mean_{s->d}(lin_l(x)[s]) + lin_r(x)`` — transform, then aggregate (sageconv_updated.py:71-76)."""
    _fn = ops.SageConvUpdatedFn
