"""LayerNorm twins of the conv layers (mirror of alignn/models/alignn_atomwise.py:127-246).

`train.py` trains ALIGNNAtomWise, whose EdgeGatedGraphConv / ALIGNNConv / MLPLayer use
`nn.LayerNorm` where alignn.py uses `nn.BatchNorm1d`; parameter names are the same.
"""
from __future__ import annotations

from torch import nn

from .conv import ALIGNNConvBase, EdgeGatedGraphConvBase


class MLPLayer(nn.Module):
    """Linear -> LayerNorm -> SiLU (alignn/models/utils.py:277-292)."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.layer = nn.Sequential(nn.Linear(in_features, out_features), nn.LayerNorm(out_features), nn.SiLU())

    def forward(self, x):
        return self.layer(x)


class EdgeGatedGraphConv(EdgeGatedGraphConvBase):
    """LayerNorm variant (alignn/models/alignn_atomwise.py:127-208)."""

    def __init__(self, input_features: int, output_features: int, residual: bool = True):
        super().__init__(input_features, output_features, residual, norm="layernorm")


class ALIGNNConv(ALIGNNConvBase):
    """alignn/models/alignn_atomwise.py:211-246."""

    conv_cls = EdgeGatedGraphConv
