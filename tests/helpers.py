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


def assert_close(a, b, tol=REL_TOL, what="", atol=2e-6, floor=0.0):
    """max|a-b| <= tol * max(max|b|, floor) + atol.  `floor` / `atol` only matter for tensors that are
    mathematically zero in the reference (e.g. the gradient of a bias that feeds a train-mode
    BatchNorm): there the error is judged against the scale of the sibling gradients."""
    a = torch.as_tensor(np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b), dtype=torch.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.numel() == 0:
        return
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    assert err <= tol * max(scale, floor) + atol, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (rel {err / max(scale, 1e-30):.3e} > {tol:.1e})"


def assert_dict_close(out, ref, tol=REL_TOL, what="", keys=None):
    """Compare every tensor of `ref`; parameter gradients ('g.*') that are ~0 in the reference are
    judged against 1% of the largest parameter-gradient magnitude of the same layer."""
    def mx(v):
        v = v.detach().cpu() if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
        return float(v.abs().max()) if v.numel() else 0.0
    gfloor = 1e-2 * max([mx(v) for k, v in ref.items() if k.startswith("g.")] or [0.0])
    for k in (keys or ref.keys()):
        if out[k] is None and ref[k] is None:
            continue
        assert_close(out[k], ref[k], tol=tol, what=f"{what} {k}", floor=gfloor if k.startswith("g.") else 0.0)
