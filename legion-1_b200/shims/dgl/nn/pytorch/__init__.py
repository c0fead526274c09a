"""SAGEConv (mean) and GraphConv (norm='both') with DGL's parameterisation and forward semantics on blocks."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class SAGEConv(nn.Module):
    """h_v = W_self h_v + W_neigh mean_{u in N(v)} h_u + b   (DGL: fc_self has no bias, one shared bias)."""

    def __init__(self, in_feats, out_feats, aggregator_type="mean", feat_drop=0.0, bias=True, norm=None, activation=None):
        super().__init__()
        assert aggregator_type == "mean", "the Legion trainers use the mean aggregator"
        self.fc_self = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_feats)) if bias else None
        self.feat_drop, self.norm, self.activation = nn.Dropout(feat_drop), norm, activation
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, block, feat):
        h_src = self.feat_drop(feat)
        h_dst = h_src[:block.number_of_dst_nodes()]
        deg = block.in_degrees().clamp(min=1).to(h_src.dtype).unsqueeze(1)
        lin_before = self.fc_neigh.in_features > self.fc_neigh.out_features      # DGL applies the linear first when it shrinks
        msg = self.fc_neigh(h_src) if lin_before else h_src
        h_neigh = block.sum_messages(msg) / deg
        # W_self h_v + b in one GEMM epilogue, W_neigh mean(h_u) accumulated into it by the second GEMM (beta = 1):
        # no separate [n_dst x out] add passes
        out = F.linear(h_dst, self.fc_self.weight, self.bias)
        out = out + h_neigh if lin_before else torch.addmm(out, h_neigh, self.fc_neigh.weight.t())
        if self.activation is not None:
            out = self.activation(out)
        if self.norm is not None:
            out = self.norm(out)
        return out


class GraphConv(nn.Module):
    """h_v = b + sum_{u in N(v)} h_u W / sqrt(d_out(u) d_in(v))   (norm='both')."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None, allow_zero_in_degree=False):
        super().__init__()
        assert norm == "both"
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        nn.init.xavier_uniform_(self.weight)
        self.bias = nn.Parameter(torch.zeros(out_feats)) if bias else None
        self.activation, self.in_feats, self.out_feats = activation, in_feats, out_feats

    def forward(self, block, feat):
        out_deg = block.out_degrees().clamp(min=1).to(feat.dtype)
        h = feat * out_deg.pow(-0.5).unsqueeze(1)
        if self.in_feats > self.out_feats:
            h = block.sum_messages(h @ self.weight)
        else:
            h = block.sum_messages(h) @ self.weight
        in_deg = block.in_degrees().clamp(min=1).to(feat.dtype)
        h = h * in_deg.pow(-0.5).unsqueeze(1)
        if self.bias is not None:
            h = h + self.bias
        return self.activation(h) if self.activation is not None else h
