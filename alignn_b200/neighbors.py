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
multi-edges to different images and self-image bonds occur.  `radius_graph_device` / `crystal_graph_device` run the
same scan (and the index / line-graph construction) on the GPU with bit-identical results.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch

from .graph import Graph, bond_cosines


def radius_graph_device(lattice_mat, cart_coords, cutoff: float = 5.0, bond_tol: float = 0.5, atol: float = 1e-5,
                        cutoff_extra: float = 0.5, device="cuda"):
    """`radius_graph` with the distance scan ON THE GPU (alignn_b200_radius_graph_offsets / _fill: one warp per atom,
    double precision with the host builder's operation order): identical bonds in identical order and identical fp32
    displacement vectors, returned as CUDA tensors (u int32, v int32, r float32 [E,3], image_index int32) plus the
    [I,3] cell table -- what an MD loop needs to rebuild g and L(g) every step without leaving the device
    (alignn/ff/calculators.py:284-291 rebuilds them on the CPU)."""
    from . import _lib
    lib = _lib.load()
    dev = torch.device(device)
    lat = np.asarray(lattice_mat, dtype=np.float64)
    X = np.asarray(cart_coords, dtype=np.float64)
    n = X.shape[0]
    frac = X @ np.linalg.inv(lat)
    Xd = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
    while True:
        recp = 2 * math.pi * np.linalg.inv(lat).T
        recp_len = np.sqrt((recp ** 2).sum(1))
        maxr = np.ceil((cutoff + bond_tol) * recp_len / (2 * math.pi))
        nmin = np.floor(frac.min(0)) - maxr
        nmax = np.ceil(frac.max(0)) + maxr
        ranges = [np.arange(a, b, dtype=np.float64) for a, b in zip(nmin, nmax)]
        cells = np.stack(np.meshgrid(*ranges, indexing="ij"), -1).reshape(-1, 3)
        sh = torch.from_numpy(np.ascontiguousarray(cells @ lat)).to(dev)
        off = torch.empty(n + 1, device=dev, dtype=torch.int32)
        nb = int(lib.alignn_b200_radius_graph_workspace_bytes(n))
        ws = torch.empty(max(nb, 1), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            _lib.check(lib.alignn_b200_radius_graph_offsets(Xd.data_ptr(), sh.data_ptr(), n, sh.shape[0], float(cutoff), float(atol),
                                                            off.data_ptr(), ws.data_ptr(), nb, st), "alignn_b200_radius_graph_offsets")
            E = int(off[-1].item())
            u, v, ci = (torch.empty(E, device=dev, dtype=torch.int32) for _ in range(3))
            r = torch.empty(E, 3, device=dev, dtype=torch.float32)
            _lib.check(lib.alignn_b200_radius_graph_fill(Xd.data_ptr(), sh.data_ptr(), n, sh.shape[0], float(cutoff), float(atol),
                                                         off.data_ptr(), u.data_ptr(), v.data_ptr(), ci.data_ptr(), r.data_ptr(), st),
                       "alignn_b200_radius_graph_fill")
        # graphs.py:347-350: the highest-numbered atom must have a bond.  Bonds are (u, c, v)-ordered, so the last bond's
        # source is the largest source; v covers the rest.
        if E and max(int(u[-1].item()), int(v.max().item())) + 1 == n:
            return u, v, r, ci, cells
        cutoff += cutoff_extra


def crystal_graph_device(lattice_mat, cart_coords, atom_features: torch.Tensor, cutoff: float = 4.0, device="cuda"):
    """(g, lg) of one periodic structure built entirely on the GPU: radius scan, sorted-CSR index, line graph and bond
    cosines (alignn/graphs.py:267-364, 544, 588-589).  Same graphs, bit for bit, as `crystal_graph(..., "radius_graph")`
    followed by `.to(device)`."""
    u, v, r, _, _ = radius_graph_device(lattice_mat, cart_coords, cutoff=cutoff, device=device)
    n = int(np.asarray(cart_coords).shape[0])
    g = Graph(u, v, n)
    g.ndata["atom_features"] = atom_features.to(r.device)
    g.edata["r"] = r
    lg = g.line_graph(shared=True)
    lg.edata["h"] = bond_cosines(r, lg)
    return g, lg


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


def crystal_graph(lattice_mat, cart_coords, atom_features: torch.Tensor, cutoff: float = 4.0,
                  neighbor_strategy: str = "radius_graph", max_neighbors: int = 12):
    """(g, lg) for one periodic structure, laid out like `Graph.atom_dgl_multigraph` output (graphs.py:472-592):
    g.ndata['atom_features'], g.edata['r'], lg = L(g) with lg.edata['h'] = bond cosines.
    `neighbor_strategy` is the reference's switch (graphs.py:497-536): "radius_graph" or "k-nearest"."""
    if neighbor_strategy == "k-nearest":
        u, v, r, _ = knn_graph(lattice_mat, cart_coords, max_neighbors=max_neighbors, cutoff=cutoff)
    elif neighbor_strategy == "radius_graph":
        u, v, r, _ = radius_graph(lattice_mat, cart_coords, cutoff=cutoff)
    else:
        raise ValueError(f"Not implemented yet: neighbor_strategy={neighbor_strategy!r}")
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


def _all_neighbors(lat, X, cutoff, atol=1e-8):
    """Every (u, v, image, distance) with 0 < |x_v + image@lat - x_u| <= cutoff (what jarvis'
    `Atoms.get_all_neighbors(r=cutoff)` lists per site), from the native scan."""
    from . import _lib
    lib = _lib.load()
    n = X.shape[0]
    recp_len = np.sqrt(((2 * math.pi * np.linalg.inv(lat).T) ** 2).sum(1))
    maxr = np.ceil(cutoff * recp_len / (2 * math.pi)) + 1
    ranges = [np.arange(-m, m + 1, dtype=np.float64) for m in maxr]
    cells = np.stack(np.meshgrid(*ranges, indexing="ij"), -1).reshape(-1, 3)
    sh = np.ascontiguousarray(cells @ lat)
    Xc = np.ascontiguousarray(X)
    p = lambda a: a.ctypes.data  # noqa: E731
    cnt = int(lib.alignn_b200_radius_graph_count_host(p(Xc), p(sh), n, sh.shape[0], float(cutoff), float(atol)))
    u, v, ci = (np.empty(cnt, dtype=np.int64) for _ in range(3))
    r = np.empty((cnt, 3), dtype=np.float32)
    _lib.check(lib.alignn_b200_radius_graph_build_host(p(Xc), p(sh), n, sh.shape[0], float(cutoff), float(atol), cnt,
                                                       p(u), p(v), p(ci), p(r)), "alignn_b200_radius_graph_build_host")
    img = cells[ci].astype(np.int64)
    d = (X[v] + img @ lat) - X[u]
    return u, v, img, np.sqrt((d ** 2).sum(1))


def knn_graph(lattice_mat, cart_coords, max_neighbors: int = 12, cutoff: float = 8.0):
    """k-nearest-neighbour crystal graph with shell completion, made undirected the reference's way
    (`nearest_neighbor_edges` + `build_undirected_edgedata`, alignn/graphs.py:155-264, `use_canonize=True`).

    Per atom: neighbours within `cutoff` sorted by distance; everything out to the distance of the k-th one is kept
    (whole shells, so the degree is >= k); if some atom has fewer than k neighbours the cutoff grows to max(a, b, c)
    or doubles (:170-186).  Each kept pair is canonised to (min id, max id, image of the second atom relative to the
    first) and every canonical bond emits BOTH directions adjacently, (u, v, d) then (v, u, -d) (:253-257).
    Self-image bonds appear with +image and -image as two canonical bonds, as in the reference.
    Bond order: canonical pairs in order of first encounter (atoms ascending, neighbours by (distance, v, image)),
    images of one pair in lexicographic order -- the reference's order inside a pair is Python-set iteration order
    and not reproducible; no model output depends on it.
    Returns (u, v, r[float32], images[int64]).
    """
    lat = np.asarray(lattice_mat, dtype=np.float64)
    X = np.asarray(cart_coords, dtype=np.float64)
    n = X.shape[0]
    abc = np.linalg.norm(lat, axis=1)
    while True:
        u, v, img, dist = _all_neighbors(lat, X, cutoff)
        counts = np.bincount(u, minlength=n)
        if counts.min() >= max_neighbors:
            break
        cutoff = float(abc.max()) if cutoff < abc.max() else 2 * cutoff
    order = np.lexsort((img[:, 2], img[:, 1], img[:, 0], v, dist, u))     # by u, then distance, then (v, image)
    u, v, img, dist = u[order], v[order], img[order], dist[order]
    start = np.concatenate([[0], np.cumsum(counts)])
    kth = dist[start[:-1] + max_neighbors - 1]                             # distance of the k-th neighbour per atom
    keep = dist <= kth[u]
    u, v, img = u[keep], v[keep], img[keep]
    swap = v < u                                                           # canonize_edge (:127-152)
    cu, cv = np.where(swap, v, u), np.where(swap, u, v)
    cimg = np.where(swap[:, None], -img, img)
    pairs = {}
    for a, b, im in zip(cu.tolist(), cv.tolist(), map(tuple, cimg.tolist())):
        pairs.setdefault((a, b), set()).add(im)
    uu, vv, rr, ii = [], [], [], []
    frac = X @ np.linalg.inv(lat)
    for (a, b), ims in pairs.items():
        for im in sorted(ims):
            d = (frac[b] + np.asarray(im, dtype=np.float64) - frac[a]) @ lat   # :245-249
            uu += [a, b]
            vv += [b, a]
            rr += [d, -d]
            ii += [im, im]
    return (np.asarray(uu, dtype=np.int64), np.asarray(vv, dtype=np.int64), np.asarray(rr, dtype=np.float32),
            np.asarray(ii, dtype=np.int64).reshape(-1, 3))
