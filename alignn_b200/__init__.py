"""alignn_b200: B200-native (sm_100a) edge-gated graph-convolution hot path of ALIGNN.

Keeps the reference's ALIGNN / ALIGNNConfig / forward((g, lg, lat)) surface
(alignn/models/alignn.py) on top of hand-written CUDA kernels behind a C-ABI
library (include/alignn_b200.h).  See DESIGN.md.
"""
__version__ = "0.1.0"

from .graph import Graph, batch, unbatch, reverse, graph, as_graph, bond_cosines  # noqa: F401
