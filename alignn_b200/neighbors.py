"""Periodic radius-graph construction on the host (crystal -> bond list with displacement vectors).

This is the input side of the hot path for real structures and for BASELINE config 4 (ALIGNN-FF on a
~1000-atom periodic supercell); the in-tree ALIGNN-FF configs use `neighbor_strategy="radius_graph"`
(`alignn/examples/sample_data_ff/config_example_atomwise.json`).  It restates the algorithm of
`alignn/graphs.py:267-364` without jarvis-tools / DGL:

  1. enumerate the periodic images needed for `cutoff (+ bond_tol)` from the reciprocal lattice lengths,
  2. distances from every atom of the home cell to every image of every atom (native scan in the library,
     `alignn_b200_radius_graph_*_host`: 1000 atoms x 27 images in tens of milliseconds),
  3. a bond u -> v for every pair with 0 < |r| <= cutoff (`atol` guards the self distance), bonds ordered by
     (u, image index, v) exactly like `torch.where` on the [N, images*N] mask,
  4. if the highest-numbered atom ended up without any bond the cutoff is increased by `cutoff_extra` and the
     search repeated (the reference tests `dgl.graph((u, v)).num_nodes() == len(atoms)`).

Both directions of a bond appear (as two separate rows, NOT adjacent: this builder emits source-major order),
multi-edges to different images and self-image bonds occur.  The device-side version of this builder is a
"next" row of SURVEY.md section 8f; this host version is what feeds `Graph` today.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch

from .graph import Graph, bond_cosines


def radius_graph(lattice_mat, cart_coords, cutoff: float = 5.0, bond_tol: float = 0.5, atol: float = 1e-5,
                 cutoff_extra: float = 0.5) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Returns (u, v, r, images): int64 [E], int64 [E], float32 [E,3] (dst image position - src position),
    float64 [E,3] (integer cell offsets of the destination image)."""
    from . import _lib
    lib = _lib.load()
    lat = np.asarray(lattice_mat, dtype=np.float64)
    X = np.asarray(cart_coords, dtype=np.float64)
    n = X.shape[0]
    frac = X @ np.linalg.inv(lat)
    while True:
        recp = 2 * math.pi * np.linalg.inv(lat).T
        recp_len = np.sqrt((recp ** 2).sum(1))
        maxr = np.ceil((cutoff + bond_tol) * recp_len / (2 * math.pi))
        nmin = np.floor(frac.min(0)) - maxr
        nmax = np.ceil(frac.max(0)) + maxr
        ranges = [np.arange(a, b, dtype=np.float64) for a, b in zip(nmin, nmax)]
        cells = np.stack(np.meshgrid(*ranges, indexing="ij"), -1).reshape(-1, 3)        # cartesian_prod order
        shifts = cells @ lat                                                             # [I, 3]
        # native scan (csrc/graph_host.cu), same double-precision arithmetic and bond order as the restatement
        Xc = np.ascontiguousarray(X)
        sh = np.ascontiguousarray(shifts)
        p = lambda a: a.ctypes.data  # noqa: E731
        cnt = int(lib.alignn_b200_radius_graph_count_host(p(Xc), p(sh), n, sh.shape[0], float(cutoff), float(atol)))
        if cnt < 0:
            raise RuntimeError("alignn_b200_radius_graph_count_host failed")
        u, v, ci = (np.empty(cnt, dtype=np.int64) for _ in range(3))
        r = np.empty((cnt, 3), dtype=np.float32)
        _lib.check(lib.alignn_b200_radius_graph_build_host(p(Xc), p(sh), n, sh.shape[0], float(cutoff), float(atol), cnt,
                                                           p(u), p(v), p(ci), p(r)), "alignn_b200_radius_graph_build_host")
        if cnt and max(int(u.max()), int(v.max())) + 1 == n:
            return u, v, r, cells[ci]
        cutoff += cutoff_extra


def crystal_graph(lattice_mat, cart_coords, atom_features: torch.Tensor, cutoff: float = 4.0):
    """(g, lg) for one periodic structure, laid out like `Graph.atom_dgl_multigraph` output (graphs.py:472-592):
    g.ndata['atom_features'], g.edata['r'], lg = L(g) with lg.edata['h'] = bond cosines."""
    u, v, r, _ = radius_graph(lattice_mat, cart_coords, cutoff=cutoff)
    g = Graph(u, v, int(np.asarray(cart_coords).shape[0]))
    g.ndata["atom_features"] = atom_features
    g.edata["r"] = torch.from_numpy(r)
    lg = g.line_graph(shared=True)
    lg.edata["h"] = bond_cosines(g.edata["r"], lg)
    return g, lg


def diamond_supercell(reps: int = 5, a: float = 5.431, jitter: float = 0.0, seed: int = 0):
    """Diamond-cubic silicon supercell, 8 * reps^3 atoms (reps = 5 -> 1000 atoms: BASELINE config 4 shape).
    Returns (lattice_mat [3,3], cart_coords [N,3])."""
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    cells = np.stack(np.meshgrid(*[np.arange(reps)] * 3, indexing="ij"), -1).reshape(-1, 3)
    frac = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3) / reps
    lat = np.eye(3) * a * reps
    X = frac @ lat
    if jitter:
        X = X + np.random.default_rng(seed).normal(scale=jitter, size=X.shape)
    return lat, X
