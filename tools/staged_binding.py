"""ctypes binding of the experimental library (alignn_b200/csrc/staged/egc_fused.h: the fully fused one-kernel conv
forward / backward), shared by tests/test_staged.py and tools/bench_fused*.py.  Not part of the product: the shipped
binding is alignn_b200/_lib.py.  (The device-side structure builders that used to live here are shipped now:
alignn_b200/csrc/graph_device.cu, include/alignn_b200.h.)"""
import ctypes as C
import os

import numpy as np

import build_staged


class FusedArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("Nn", C.c_int64), ("Ne", C.c_int64), ("d", C.c_int32),
                ("norm_edges", C.c_int32), ("residual", C.c_int32), ("epilogue_groups", C.c_int32),
                ("gate_eps", C.c_float), ("ln_eps", C.c_float),
                ("y", C.c_void_p), ("w_image", C.c_void_p), ("bias", C.c_void_p), ("P", C.c_void_p),
                ("src", C.c_void_p), ("dst", C.c_void_p), ("in_ptr", C.c_void_p), ("in_eid", C.c_void_p),
                ("tiles", C.c_void_p), ("num_tiles", C.c_int32), ("e_w", C.c_void_p), ("e_b", C.c_void_p),
                ("M", C.c_void_p), ("y_out", C.c_void_p), ("XP", C.c_void_p), ("S", C.c_void_p), ("H", C.c_void_p),
                ("partials", C.c_void_p), ("stream", C.c_void_p)]


def load():
    """Build if stale (needs nvcc) or fall back to the prebuilt .so that travelled with the tree."""
    try:
        path = build_staged.build()
    except Exception:
        if not os.path.exists(build_staged.LIB):
            raise
        path = build_staged.LIB
    lib = C.CDLL(path)
    lib.alignn_b200_segment_tiles_host.restype = C.c_int64
    lib.alignn_b200_segment_tiles_host.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    lib.alignn_b200_egc_forward_fused.restype = C.c_int
    lib.alignn_b200_egc_forward_fused.argtypes = [C.POINTER(FusedArgs)]
    lib.alignn_b200_egc_fused_partial_rows.restype = C.c_int
    lib.alignn_b200_egc_fused_partial_rows.argtypes = [C.c_int32]
    lib.alignn_b200_staged_last_cuda_error.restype = C.c_int
    return lib


def pack_tiles(lib, in_ptr: np.ndarray):
    """(num_tiles, int32 [num_tiles,4] descriptors {v0, nseg, p0, rows}); num_tiles < 0 on failure (-2: a node has
    more than 128 in-edges)."""
    in_ptr = np.ascontiguousarray(in_ptr, dtype=np.int32)
    n = int(lib.alignn_b200_segment_tiles_host(in_ptr.ctypes.data, in_ptr.size - 1, None, 0))
    if n < 0:
        return n, None
    tiles = np.zeros((max(n, 1), 4), dtype=np.int32)
    assert lib.alignn_b200_segment_tiles_host(in_ptr.ctypes.data, in_ptr.size - 1, tiles.ctypes.data, n) == n
    return n, tiles[:n]


def fused_forward(lib, ix, tiles_d, n_tiles, y, img, b_eg, P, norm_edges, train, e_w=None, e_b=None, residual=True,
                  out=None, need_edge_out=True, groups=1):
    """Launch the fused kernel on torch's current stream.  `out` (a dict from a previous call) is reused if given."""
    import torch
    from alignn_b200 import ops
    from alignn_b200._lib import ptr, stream_ptr
    dev = y.device if y.numel() else P.device
    Nn, d = P.shape[0], P.shape[1] // 4
    Ne = y.shape[0]
    if out is None:
        new = lambda *s: torch.full(s, float("nan"), device=dev, dtype=torch.float32)  # noqa: E731
        rows = lib.alignn_b200_egc_fused_partial_rows(n_tiles)
        out = dict(M=new(Ne, d) if train else None, XP=new(Nn, d), S=new(Nn, d) if train else None,
                   H=new(Nn, d) if train else None, partials=new(rows, 2, d) if norm_edges == ops.NORM_STATS else None,
                   y_out=new(Ne, d) if (norm_edges != ops.NORM_STATS and need_edge_out) else None)
    a = FusedArgs(struct_size=C.sizeof(FusedArgs), Nn=Nn, Ne=Ne, d=d, norm_edges=norm_edges, residual=int(residual),
                  epilogue_groups=groups, gate_eps=1e-6, ln_eps=1e-5, y=ptr(y), w_image=ops.ptr_any(img.buf), bias=ptr(b_eg), P=ptr(P),
                  src=ptr(ix.src), dst=ptr(ix.dst), in_ptr=ptr(ix.in_ptr), in_eid=None if ix.dst_sorted else ptr(ix.in_eid),
                  tiles=ptr(tiles_d), num_tiles=n_tiles, e_w=ptr(e_w), e_b=ptr(e_b), M=ptr(out["M"]), y_out=ptr(out["y_out"]),
                  XP=ptr(out["XP"]), S=ptr(out["S"]), H=ptr(out["H"]), partials=ptr(out["partials"]), stream=stream_ptr())
    rc = lib.alignn_b200_egc_forward_fused(C.byref(a))
    if rc != 0:
        raise RuntimeError(f"alignn_b200_egc_forward_fused -> {rc} (cuda error {lib.alignn_b200_staged_last_cuda_error()})")
    return out


def conv_forward_like(lib, ix, tiles_d, n_tiles, x, y, img, b_eg, P, n_w, n_b, e_w, e_b, *, norm_nodes, norm_edges,
                      residual=True, save=True, need_edge_out=True, groups=1):
    """The whole post-Linear forward of one conv built from the fused kernel + node tail, with the output contract of
    `alignn_b200.ops.egc_forward` (x_out / y_out are None in STATS mode; partials_e / partials_n then feed
    `ops.bn_finalize(..., which=0, ...)`).  This is what conv.py will call once the kernel is validated."""
    import torch
    from alignn_b200 import _lib, ops
    from alignn_b200._lib import ptr, stream_ptr
    main = _lib.load()
    Nn, d = x.shape
    stats = norm_nodes == ops.NORM_STATS or norm_edges == ops.NORM_STATS
    train = save or stats
    out = fused_forward(lib, ix, tiles_d, n_tiles, y, img, b_eg, P, norm_edges, train, e_w, e_b, residual,
                        need_edge_out=need_edge_out, groups=groups)
    out["partials_e"] = out.pop("partials")
    out["partials_n"] = None
    out["x_out"] = None
    XP = out["XP"]
    if norm_nodes == ops.NORM_STATS:
        rows = ops.partial_rows(Nn, d)
        part = torch.empty(rows, 2, d, device=x.device, dtype=torch.float32)
        _lib.check(main.alignn_b200_rowstats_partials(ptr(XP), Nn, d, ptr(part), rows, stream_ptr()), "rowstats_partials")
        out["partials_n"] = part
    elif norm_nodes == ops.NORM_AFFINE:
        out["x_out"] = ops.affine_silu_residual(XP, x if residual else None, n_w, n_b)
    else:
        xo = torch.empty_like(XP)
        lib.alignn_b200_ln_silu_residual.restype = C.c_int
        lib.alignn_b200_ln_silu_residual.argtypes = [C.c_void_p] * 4 + [C.c_float, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        rc = lib.alignn_b200_ln_silu_residual(ptr(XP), ptr(x) if residual else None, ptr(n_w), ptr(n_b), 1e-5, ptr(xo), Nn, d,
                                              stream_ptr())
        if rc != 0:
            raise RuntimeError(f"alignn_b200_ln_silu_residual -> {rc}")
        out["x_out"] = xo
    return out


# ---- fused backward (csrc/staged/egc_bwd_fused_tc.cu + node side in node_tail.cu) ----------------------------------
class BwdFusedArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("Nn", C.c_int64), ("Ne", C.c_int64), ("d", C.c_int32), ("residual", C.c_int32),
                ("M", C.c_void_p), ("gy_out", C.c_void_p), ("P", C.c_void_p), ("GSh", C.c_void_p), ("GS", C.c_void_p),
                ("w_image", C.c_void_p), ("src", C.c_void_p), ("dst", C.c_void_p), ("in_ptr", C.c_void_p), ("in_eid", C.c_void_p),
                ("tiles", C.c_void_p), ("num_tiles", C.c_int32),
                ("e_w", C.c_void_p), ("e_b", C.c_void_p), ("e_mean", C.c_void_p), ("e_rstd", C.c_void_p), ("e_c1", C.c_void_p),
                ("e_c2", C.c_void_p), ("GM", C.c_void_p), ("gy", C.c_void_p), ("GPB", C.c_void_p), ("ld_gpb", C.c_int64),
                ("partials", C.c_void_p), ("stream", C.c_void_p)]


def backward_fused(lib, ix, tiles_d, n_tiles, P, M, XP, S, H, gx_out, gy_out, n, e, W_eg_T_img, residual=True, need_gy=True):
    """Train-mode BatchNorm backward of one conv up to (GM, GP[:, 2d:4d], gy, column sums): node kernel + fused edge
    kernel.  n / e: dicts with w, b, mean, rstd, c1, c2 (as for `ops.egc_backward`).  GP[:, 0:2d] (the source-keyed
    sums) is NOT produced here: it still comes from egc_backward_src_kernel."""
    import torch
    from alignn_b200 import ops
    from alignn_b200._lib import ptr, stream_ptr
    lib.alignn_b200_egc_backward_nodes.restype = C.c_int
    lib.alignn_b200_egc_backward_nodes.argtypes = [C.c_void_p] * 10 + [C.c_float, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.alignn_b200_egc_backward_fused.restype = C.c_int
    lib.alignn_b200_egc_backward_fused.argtypes = [C.POINTER(BwdFusedArgs)]
    Nn, d = XP.shape
    Ne = M.shape[0]
    dev = XP.device
    new = lambda *s: torch.full(s, float("nan"), device=dev, dtype=torch.float32)  # noqa: E731
    GP, GSh, GS, GM = new(Nn, 4 * d), new(Nn, d), new(Nn, d), new(Ne, d)
    gy = new(Ne, d) if need_gy else None
    rows_n = ops.partial_rows(Nn, d)
    part_n = new(rows_n, d)
    rc = lib.alignn_b200_egc_backward_nodes(ptr(XP), ptr(gx_out), ptr(S), ptr(H), ptr(n["w"]), ptr(n["b"]), ptr(n["mean"]),
                                            ptr(n["rstd"]), ptr(n["c1"]), ptr(n["c2"]), 1e-6, Nn, d,
                                            GP.data_ptr() + 3 * d * 4, 4 * d, ptr(GSh), ptr(GS), ptr(part_n), rows_n, stream_ptr())
    if rc != 0:
        raise RuntimeError(f"alignn_b200_egc_backward_nodes -> {rc}")
    rows_e = lib.alignn_b200_egc_fused_partial_rows(n_tiles)
    part_e = new(rows_e, d)
    g = lambda k: ptr(e.get(k)) if gy_out is not None else None  # noqa: E731
    a = BwdFusedArgs(struct_size=C.sizeof(BwdFusedArgs), Nn=Nn, Ne=Ne, d=d, residual=int(residual), M=ptr(M), gy_out=ptr(gy_out),
                     P=ptr(P), GSh=ptr(GSh), GS=ptr(GS), w_image=ops.ptr_any(W_eg_T_img.buf), src=ptr(ix.src), dst=ptr(ix.dst),
                     in_ptr=ptr(ix.in_ptr), in_eid=None if ix.dst_sorted else ptr(ix.in_eid), tiles=ptr(tiles_d),
                     num_tiles=n_tiles, e_w=g("w"), e_b=g("b"), e_mean=g("mean"), e_rstd=g("rstd"), e_c1=g("c1"), e_c2=g("c2"),
                     GM=ptr(GM), gy=ptr(gy), GPB=GP.data_ptr() + 2 * d * 4, ld_gpb=4 * d, partials=ptr(part_e), stream=stream_ptr())
    rc = lib.alignn_b200_egc_backward_fused(C.byref(a))
    if rc != 0:
        raise RuntimeError(f"alignn_b200_egc_backward_fused -> {rc} (cuda error {lib.alignn_b200_staged_last_cuda_error()})")
    return dict(GM=GM, GP=GP, GSh=GSh, gy=gy, sum_gD=part_n.sum(0), sum_gm=part_e.sum(0))
