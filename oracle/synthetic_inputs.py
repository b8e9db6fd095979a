"""TEST / BASELINE INFRASTRUCTURE: the synthetic JARVIS-shaped batch of `alignn_b200.synthetic.make_batch(regular=True)`
rebuilt with oracle types only, so that the CPU reference arm of bench.py (`--impl reference`, `cpu_baseline`) never
imports the product package or loads its shared library.  Same generator, same draws: tests/test_oracle_golden.py
checks that both produce identical graphs and features."""
import numpy as np
import torch

from . import alignn_oracle as O


def _atom_features(rng, n, nfeat):
    species = rng.integers(0, 20, size=n)
    table = (np.random.default_rng(7).random((20, nfeat)) < 0.12).astype(np.float32)
    return table[species]


def _regular_crystal(rng, n, k):
    half = k // 2
    offs = rng.choice(np.arange(1, max(n, 2)), size=half, replace=(n - 1 < half)) if n > 1 else np.zeros(half, int)
    u = np.repeat(np.arange(n), half)
    v = (u + np.tile(offs, n)) % n
    length = rng.uniform(1.5, 6.0, size=u.shape[0])
    dirs = rng.normal(size=(u.shape[0], 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rv = (dirs * length[:, None]).astype(np.float32)
    src = np.stack([u, v], 1).reshape(-1)           # (u,v) then (v,u) adjacent, alignn/graphs.py:253-257
    dst = np.stack([v, u], 1).reshape(-1)
    r = np.stack([rv, -rv], 1).reshape(-1, 3)
    return src, dst, r


def make_batch(batch_size=64, atoms=30, k=12, seed=123, atom_input_features=92):
    """-> (g, lg, lattice [B,3,3], target [B]) as oracle OGraphs; lg.edata['h'] = bond cosines (alignn/graphs.py:847-864)."""
    rng = np.random.default_rng(seed)
    gs, lats, ys = [], [], []
    for _ in range(batch_size):
        src, dst, r = _regular_crystal(rng, atoms, k)
        g = O.OGraph(src, dst, atoms)
        g.ndata["atom_features"] = torch.from_numpy(_atom_features(rng, atoms, atom_input_features))
        g.edata["r"] = torch.from_numpy(np.ascontiguousarray(r, dtype=np.float32))
        gs.append(g)
        lats.append(torch.eye(3, dtype=torch.float32) * (atoms * 18.0) ** (1.0 / 3.0))
        ys.append(float(rng.normal()))
    g = O.batch(gs)
    lg = O.line_graph(g)
    lg.edata["h"] = O.bond_cosines(g.edata["r"], lg.src, lg.dst)
    return g, lg, torch.stack(lats), torch.tensor(ys, dtype=torch.float32)
