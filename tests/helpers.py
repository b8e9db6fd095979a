"""Shared test helpers: conversions between the product Graph and the oracle's OGraph, tolerances."""
import numpy as np
import torch

from oracle import alignn_oracle as O

# north_star: "outputs match the reference DGL path ... within 1e-4 rel fp32"
REL_TOL = 1e-4


def to_oracle(g, dtype=None):
    s, d = g.edges()
    og = O.OGraph(s.cpu().long(), d.cpu().long(), g.num_nodes(), g.batch_num_nodes(), g.batch_num_edges())
    for src, dst in ((g.ndata, og.ndata), (g.edata, og.edata)):
        for k, v in src.items():
            v = v.detach().cpu()
            dst[k] = v.to(dtype) if (dtype is not None and v.is_floating_point()) else v
    return og


def rel_err(a, b):
    """max |a-b| / max(|b|_max, 1e-30): relative to the tensor's scale (fp32 parity metric)."""
    a = torch.as_tensor(np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b), dtype=torch.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_close(a, b, tol=REL_TOL, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"
