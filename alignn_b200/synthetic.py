"""Seeded synthetic JARVIS-DFT-shaped crystal graph batches.

There is no network for datasets, so benchmark and parity inputs are synthetic but
keep the reference's input contract (alignn/graphs.py:230-264,
472-592): a directed multigraph per crystal where both directions of every bond are
present and adjacent in edge order, bond vectors `r` (src -> dst displacement, the
reverse edge carries -r), CGCNN-style 92-d atom features, a [3,3] lattice and a
scalar label per crystal.

Two layouts:
  * regular=True  -- every atom has exactly k in- and k out-bonds (a periodic
    "circulant" lattice).  Gives the headline shapes exactly: B=64, n=30, k=12
    => N=1920, E=23 040, T=E*k=276 480.
  * regular=False -- atoms at random positions in a periodic cubic cell, true
    k-nearest-neighbour bonds over periodic images, symmetrised the way
    build_undirected_edgedata does (alignn/graphs.py:230-264); in-degree >= k and
    not constant, multi-edges and self-image bonds occur for small cells.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .graph import Graph, batch as batch_graphs, bond_cosines


def _atom_features(rng: np.random.Generator, n: int, nfeat: int) -> np.ndarray:
    """One-hot-ish 0/1 rows like the CGCNN atom table (alignn/config.py:23)."""
    species = rng.integers(0, 20, size=n)
    table = (np.random.default_rng(7).random((20, nfeat)) < 0.12).astype(np.float32)
    return table[species]


def _regular_crystal(rng, n, k):
    half = k // 2
    assert k % 2 == 0
    offs = rng.choice(np.arange(1, max(n, 2)), size=half, replace=(n - 1 < half)) if n > 1 else np.zeros(half, int)
    u = np.repeat(np.arange(n), half)
    v = (u + np.tile(offs, n)) % n
    length = rng.uniform(1.5, 6.0, size=u.shape[0])
    dirs = rng.normal(size=(u.shape[0], 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rv = (dirs * length[:, None]).astype(np.float32)
    src = np.stack([u, v], 1).reshape(-1)           # (u,v) then (v,u) adjacent, graphs.py:253-257
    dst = np.stack([v, u], 1).reshape(-1)
    r = np.stack([rv, -rv], 1).reshape(-1, 3)
    return src, dst, r


def _knn_crystal(rng, n, k):
    a = (n * 18.0) ** (1.0 / 3.0)                   # ~18 A^3 per atom
    frac = rng.random((n, 3))
    pos = frac * a
    rep = 1
    while (2 * rep + 1) ** 3 * n < 4 * k + 1:
        rep += 1
    rng_img = np.arange(-rep, rep + 1)
    imgs = np.stack(np.meshgrid(rng_img, rng_img, rng_img, indexing="ij"), -1).reshape(-1, 3)
    # displacement from atom i to every image of atom j
    disp = pos[None, :, None, :] + imgs[None, None, :, :] * a - pos[:, None, None, :]   # [n, n, I, 3]
    dist = np.linalg.norm(disp, axis=-1)
    pairs = {}
    for i in range(n):
        d = dist[i].reshape(-1)
        d_valid = np.where(d > 1e-8, d, np.inf)
        idx = np.argsort(d_valid, kind="stable")[:k]
        for f in idx:
            j, im = divmod(int(f), imgs.shape[0])
            image = imgs[im]
            # canonical orientation, as in nearest_neighbor_edges (graphs.py:155-227)
            if j < i:
                key = (j, i, tuple((-image).tolist()))
            elif j == i:
                key = (i, i, tuple(image.tolist()))
                neg = (i, i, tuple((-image).tolist()))
                if neg in pairs:
                    continue
            else:
                key = (i, j, tuple(image.tolist()))
            pairs[key] = True
    src, dst, r = [], [], []
    for (u, v, image) in pairs:
        dvec = pos[v] + np.asarray(image) * a - pos[u]
        src += [u, v]
        dst += [v, u]
        r += [dvec, -dvec]
    return np.asarray(src), np.asarray(dst), np.asarray(r, dtype=np.float32)


def make_crystal(rng: np.random.Generator, n: int, k: int = 12, regular: bool = True,
                 atom_input_features: int = 92) -> Tuple[Graph, torch.Tensor, float]:
    src, dst, r = (_regular_crystal if regular else _knn_crystal)(rng, n, k)
    g = Graph(src, dst, n)
    g.ndata["atom_features"] = torch.from_numpy(_atom_features(rng, n, atom_input_features))
    g.edata["r"] = torch.from_numpy(np.ascontiguousarray(r, dtype=np.float32))
    a = (n * 18.0) ** (1.0 / 3.0)
    lat = torch.eye(3, dtype=torch.float32) * a
    return g, lat, float(rng.normal())


def make_batch(batch_size: int = 64, atoms: int = 30, k: int = 12, seed: int = 123,
               regular: bool = True, vary_atoms: bool = False,
               atom_input_features: int = 92):
    """Returns (g, lg, lattice[B,3,3], target[B]) on the CPU, collated like
    collate_line_graph (alignn/lmdb_dataset.py:87-108): lg.edata['h'] = bond cosines."""
    rng = np.random.default_rng(seed)
    gs: List[Graph] = []
    lats, ys = [], []
    for _ in range(batch_size):
        n = atoms
        if vary_atoms:
            n = int(np.clip(np.round(rng.lognormal(np.log(atoms) - 0.125, 0.5)), 2, 100))
        g, lat, y = make_crystal(rng, n, k, regular, atom_input_features)
        gs.append(g)
        lats.append(lat)
        ys.append(y)
    g = batch_graphs(gs)
    lg = g.line_graph(shared=True)
    lg.edata["h"] = bond_cosines(g.edata["r"], lg)
    return g, lg, torch.stack(lats), torch.tensor(ys, dtype=torch.float32)


def make_segment_sweep(num_edges: int, d: int = 256, fan_in: int = 12, seed: int = 123):
    """BASELINE.json config 5: Nn = Ne/fan_in nodes, dst = repeat_interleave(arange(Nn), fan_in),
    uniform random src; sigma ~ U(0,1), Bh ~ N(0,1)."""
    gen = torch.Generator().manual_seed(seed)
    nn_ = max(1, num_edges // fan_in)
    ne = nn_ * fan_in
    dst = torch.repeat_interleave(torch.arange(nn_), fan_in)
    src = torch.randint(0, nn_, (ne,), generator=gen)
    g = Graph(src, dst, nn_)
    sigma = torch.rand(ne, d, generator=gen)
    bh = torch.randn(nn_, d, generator=gen)
    return g, bh, sigma
