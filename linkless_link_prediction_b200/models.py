"""Model library with the reference's class / constructor / forward / state-dict surface
(``src/models.py``): ``MLP`` (:6-54), ``SAGE`` (:82-119), ``LinkPredictor`` (:121-150).
``GCN`` (:56-80) is out of scope (never selected by the reference scripts).

Every dense layer runs through the fused GEMM epilogues (bias + relu + dropout) of
``libllp_b200.so``; activations live in the compute dtype chosen by ``ops.set_compute_dtype``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .sageconv import SAGEConv, _SageBase


def _linear(x, lin: nn.Linear, relu: bool, p: float, in_gate: float = 0.0, defer_gate: bool = False):
    seed, offset = ops._dropout_seed() if p > 0 else (0, 0)
    return ops.LinearFn.apply(x, lin.weight, lin.bias, relu, p, seed, offset, in_gate, defer_gate)


FUSE_EDGE_MLP = True  # LinkPredictor.score: gather + lin1 + head as one tcgen05 kernel when the shapes allow it


def _append_norm(norms: nn.ModuleList, norm_type: str, width: int) -> None:
    """The norm layer the reference appends next to a hidden layer (models.py:27-37,90-100)."""
    if norm_type == "batch":
        norms.append(nn.BatchNorm1d(width))
    elif norm_type == "layer":
        norms.append(nn.LayerNorm(width))


def _gate_scale(p: float) -> float:
    """Backward factor of relu + dropout(p) taken from the saved output: 1/(1-p) where the output is positive."""
    return 1.0 / (1.0 - p)


class MLP(nn.Module):
    def __init__(self, num_layers, input_dim, hidden_dim, output_dim, dropout_ratio, norm_type="none"):
        super().__init__()
        self.num_layers, self.norm_type = num_layers, norm_type
        self.dropout = nn.Dropout(dropout_ratio)
        self.layers, self.norms = nn.ModuleList(), nn.ModuleList()
        if num_layers == 1:
            self.layers.append(nn.Linear(input_dim, output_dim))
        else:
            self.layers.append(nn.Linear(input_dim, hidden_dim))
            _append_norm(self.norms, norm_type, hidden_dim)
            for _ in range(num_layers - 2):
                self.layers.append(nn.Linear(hidden_dim, hidden_dim))
                _append_norm(self.norms, norm_type, hidden_dim)
            self.layers.append(nn.Linear(hidden_dim, output_dim))

    def reset_parameters(self):
        for layer in self.layers:
            layer.reset_parameters()

    def forward(self, feats):
        h = ops.to_compute(feats, cache=True)
        p = float(self.dropout.p) if self.training else 0.0
        if self.norm_type != "none":
            # models.py:48-53 with a norm layer between the linear layer and relu: the GEMM runs bare and the norm / relu /
            # dropout go through torch (the reference drivers never enable norms, so this branch is off the measured path)
            for l, layer in enumerate(self.layers):
                h = _linear(h, layer, relu=False, p=0.0)
                if l != self.num_layers - 1:
                    h = self.dropout(F.relu(self.norms[l](h.float()))).to(ops.compute_dtype()).contiguous()
            return h
        for l, layer in enumerate(self.layers):
            last = l == self.num_layers - 1
            # every hidden activation is consumed by the next layer, which applies its relu/dropout mask in backward
            h = _linear(h, layer, relu=not last, p=0.0 if last else p, in_gate=_gate_scale(p) if l > 0 else 0.0,
                        defer_gate=not last)
        return h


class SAGE(nn.Module):
    def __init__(self, data_name, in_channels, hidden_channels, out_channels, num_layers, dropout, conv_layer=SAGEConv,
                 norm_type="none"):
        super().__init__()
        self.convs, self.norms, self.norm_type = nn.ModuleList(), nn.ModuleList(), norm_type
        _append_norm(self.norms, norm_type, hidden_channels)
        self.convs.append(conv_layer(in_channels, hidden_channels))
        for _ in range(num_layers - 2):
            self.convs.append(conv_layer(hidden_channels, hidden_channels))
            _append_norm(self.norms, norm_type, hidden_channels)
        self.convs.append(conv_layer(hidden_channels, out_channels))
        self.dropout = dropout

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def forward(self, x, adj_t):
        graph = adj_t if isinstance(adj_t, ops.Graph) else ops.graph_of(adj_t, x.size(0))
        x = ops.to_compute(x, cache=True)
        p = float(self.dropout) if self.training else 0.0
        n = len(self.convs)
        deferred = False   # did the previous layer leave its relu/dropout backward mask to this one?
        for l, conv in enumerate(self.convs):
            if not isinstance(conv, _SageBase):
                raise RuntimeError("SAGE expects the SAGEConv / SAGEConv_updated layers of this package")
            last = l == n - 1
            fused = self.norm_type == "none" and not conv.normalize
            if fused:   # relu + dropout inside the layer's GEMM epilogue, their backward mask in the next layer's
                x = conv(x, graph, _relu=not last, _dropout=0.0 if last else self.dropout,
                         _in_gate=_gate_scale(p) if deferred else 0.0, _defer_gate=not last, _input_layer=l == 0)
                deferred = not last
            else:       # a norm layer (models.py:113-116) or L2 normalisation (sageconv_updated.py:78-79) sits between the
                x = conv(x, graph, _in_gate=_gate_scale(p) if deferred else 0.0, _input_layer=l == 0)   # layer and relu: those go through torch
                deferred = False
                if not last:
                    if self.norm_type != "none":
                        x = self.norms[l](x.float())
                    x = F.dropout(F.relu(x), p=self.dropout, training=self.training).to(ops.compute_dtype()).contiguous()
        return x


class LinkPredictor(nn.Module):
    def __init__(self, predictor, in_channels, hidden_channels, out_channels, num_layers, dropout):
        super().__init__()
        if predictor not in ("mlp", "inner"):
            raise ValueError(predictor)
        self.predictor = predictor
        self.lins = nn.ModuleList()
        self.lins.append(nn.Linear(in_channels, hidden_channels))
        for _ in range(num_layers - 2):
            self.lins.append(nn.Linear(hidden_channels, hidden_channels))
        self.lins.append(nn.Linear(hidden_channels, out_channels))
        self.dropout = dropout

    def reset_parameters(self):
        for lin in self.lins:
            lin.reset_parameters()

    def _head(self, z):
        """z = x_i * x_j rows [M, C] in the compute dtype -> sigmoid scores."""
        p = float(self.dropout) if self.training else 0.0
        if self.predictor == "mlp":
            hidden = len(self.lins) - 1
            for l, lin in enumerate(self.lins[:-1]):
                z = _linear(z, lin, relu=True, p=p, in_gate=_gate_scale(p) if l > 0 else 0.0, defer_gate=True)
            last = self.lins[-1]
            gate_in = _gate_scale(p) if hidden > 0 else 0.0
            if last.out_features == 1:
                return ops.ScoreHeadFn.apply(z, last.weight, last.bias, gate_in).unsqueeze(-1)
            return torch.sigmoid(_linear(z, last, relu=False, p=0.0, in_gate=gate_in).float())
        ones = torch.ones(1, z.size(1), dtype=torch.float32, device=z.device)
        return ops.ScoreHeadFn.apply(z, ones, None)  # 'inner': sigmoid(sum(x_i * x_j))

    def forward(self, x_i, x_j):
        """Reference signature (models.py:139): pre-gathered rows, 2-D ``[M,C]`` or 3-D ``[B,K,C]`` (main.py:186)."""
        lead = x_i.shape[:-1]
        z = ops.to_compute((x_i * x_j).reshape(-1, x_i.size(-1)))
        out = self._head(z)
        return out.reshape(*lead, 1) if self.predictor == "mlp" else out.reshape(*lead)

    def score(self, h, u, v, plan=None):
        """Fused path used by the step functions: scores of the edges ``(u[m], v[m])`` straight from the node
        embedding matrix ``h`` — same value as ``forward(h[u], h[v])`` without materialising the gathers.
        ``plan``: optional ``ops.EdgePlan`` of (u, v) built ahead of time for the backward."""
        lead = u.shape
        hc, uf, vf = ops.to_compute(h), u.reshape(-1).contiguous(), v.reshape(-1).contiguous()
        if self.predictor == "mlp" and len(self.lins) == 2 and self.lins[1].out_features == 1 and FUSE_EDGE_MLP \
                and ops.edge_mlp_supported(hc, self.lins[0].in_features, self.lins[0].out_features):
            # one kernel: gather-Hadamard -> lin1 (+bias, relu, dropout) -> lin2 -> sigmoid
            p = float(self.dropout) if self.training else 0.0
            _, site = ops._dropout_seed() if p > 0 else (0, 0)
            l1, l2 = self.lins[0], self.lins[1]
            out = ops.EdgeMlpFn.apply(hc, uf, vf, l1.weight, l1.bias, l2.weight, l2.bias, p, site, plan)
            return out.reshape(*lead, 1)
        z = ops.HadamardFn.apply(hc, uf, vf, plan)
        out = self._head(z)
        return out.reshape(*lead, 1) if self.predictor == "mlp" else out.reshape(*lead)
