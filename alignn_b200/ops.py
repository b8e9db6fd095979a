"""Tensor-level wrappers over the C ABI (include/alignn_b200.h).  Device memory, streams
and the autograd glue are PyTorch plumbing; all arithmetic of the edge-gated conv stage
runs in libalignn_b200.so.  Every function raises if the tensors are not fp32/int32 CUDA
tensors or if the library reports an error -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import weakref
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import NORM_AFFINE, NORM_LAYER, NORM_STATS, ptr, require_cuda, stream_ptr  # noqa: F401
from .graph import EdgeIndex


def _on_tensor_device(fn):
    """Launch on the device the operands live on, whatever the current device is (a model on cuda:1 in a process whose
    current device is cuda:0 would otherwise enqueue on the wrong device and stream)."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        dev = None
        for a in args:
            if isinstance(a, EdgeIndex):
                a = a.src
            elif isinstance(a, (list, tuple)) and a and isinstance(a[0], (list, tuple)) and a[0]:
                a = a[0][0]
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)
    return wrapper


class KernelTimer:
    """Optional CUDA-event timing of individual library calls (used by bench.py for the roofline
    of the dominant kernel: events are recorded on the launching stream, inside the timed region)."""

    def __init__(self, min_edges: int = 0):
        self.min_edges = min_edges
        self.records = {}      # name -> list of (start_event, end_event, algorithmic_bytes)

    def span(self, name: str, nbytes: int):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.records.setdefault(name, []).append((s, e, nbytes))
        return s, e

    def summary(self):
        """name -> launches, total_ms, total algorithmic bytes; plus the same for the largest launches alone
        (`big_*`: the L(g)-sized calls, bytes >= half of the largest)."""
        out = {}
        for name, rec in self.records.items():
            ms = [s.elapsed_time(e) for s, e, _ in rec]
            by = [b for _, _, b in rec]
            top = max(by)
            big = [(m, b) for m, b in zip(ms, by) if 2 * b >= top]
            out[name] = dict(launches=len(ms), total_ms=sum(ms), total_bytes=sum(by), big_launches=len(big),
                             big_ms=sum(m for m, _ in big), big_bytes=sum(b for _, b in big))
        return out


TIMER: Optional[KernelTimer] = None


class _span:
    """`with _span(name, nbytes): <library call>` -- CUDA events around the call on the launching stream when a
    KernelTimer is installed (bench.py); free otherwise."""
    __slots__ = ("ev",)

    def __init__(self, name: str, nbytes: int):
        self.ev = TIMER.span(name, int(nbytes)) if TIMER is not None else None

    def __enter__(self):
        if self.ev is not None:
            self.ev[0].record()

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()


def _check_d(d: int) -> None:
    if d not in _lib.SUPPORTED_D:
        raise RuntimeError(f"alignn_b200: unsupported feature width {d}; supported: {_lib.SUPPORTED_D}")


def partial_rows(num_nodes: int, d: int) -> int:
    return int(_lib.load().alignn_b200_egc_partial_rows(num_nodes, d))


@_on_tensor_device
def egc_forward(ix: EdgeIndex, x, y, G, P, n_w, n_b, e_w, e_b, *, norm_nodes: int, norm_edges: int,
                residual: bool, save: bool, need_edge_out: bool, gate_eps: float = 1e-6, ln_eps: float = 1e-5,
                gate_is_m: bool = False):
    """Everything of EdgeGatedGraphConv.forward after the Linear layers (alignn.py:100-127).

    gate_is_m: G already holds the pre-activation gate m (from `gemm_gather` with the e_src / e_dst addends); the
    kernel then only applies the edge norm / SiLU / residual and reduces the gated messages (second pass over the edges).

    Returns dict(x_out, y_out, M, XP, S, H, partials)."""
    lib = _lib.load()
    Nn, d = x.shape
    Ne = y.shape[0]
    _check_d(d)
    require_cuda(x, y, G, P, n_w, n_b, e_w, e_b, ix.src, ix.in_ptr, ix.in_eid)
    stats = norm_nodes == NORM_STATS or norm_edges == NORM_STATS
    save = save or stats
    dev = x.device
    new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
    x_out = None if norm_nodes == NORM_STATS else new(Nn, d)
    y_out = new(Ne, d) if (need_edge_out and norm_edges != NORM_STATS) else None
    if gate_is_m and norm_edges == NORM_STATS and need_edge_out:
        raise RuntimeError("egc_forward(gate_is_m=True): finalize the batch statistics of m first and pass NORM_AFFINE")
    M = (G if gate_is_m else new(Ne, d)) if save else None
    XP, S, H = (new(Nn, d), new(Nn, d), new(Nn, d)) if save else (None, None, None)
    partials = new(partial_rows(Nn, d), 4, d) if stats else None
    a = _lib.EgcFwdArgs(
        struct_size=C.sizeof(_lib.EgcFwdArgs), Nn=Nn, Ne=Ne, d=d, norm_nodes=norm_nodes, norm_edges=norm_edges,
        residual=int(residual), gate_is_m=int(gate_is_m), gate_eps=gate_eps, ln_eps=ln_eps,
        x=ptr(x), y=ptr(y), G=ptr(G), P=ptr(P), src=ptr(ix.src), in_ptr=ptr(ix.in_ptr),
        in_eid=None if ix.dst_sorted else ptr(ix.in_eid),
        n_w=ptr(n_w), n_b=ptr(n_b), e_w=ptr(e_w), e_b=ptr(e_b),
        x_out=ptr(x_out), y_out=ptr(y_out), M=None if gate_is_m else ptr(M), XP=ptr(XP), S=ptr(S), H=ptr(H),
        partials=ptr(partials),
        stream=stream_ptr())
    # compulsory bytes of THIS kernel: read G (or m), P (each element once), indices; residual rows; what it writes
    nb = 4 * d * (Ne + 4 * Nn) + 4 * Ne + 4 * (Nn + 1)
    nb += 4 * d * Ne * (M is not None and not gate_is_m) + 4 * d * Nn * 3 * (XP is not None)
    nb += 4 * d * Ne * (1 + int(residual)) * (y_out is not None) + 4 * d * Nn * (1 + int(residual)) * (x_out is not None)
    with _span("egc_forward" + ("<gate_is_m>" if gate_is_m else ""), nb):
        _lib.check(lib.alignn_b200_egc_forward(C.byref(a)), "alignn_b200_egc_forward")
    return dict(x_out=x_out, y_out=y_out, M=M, XP=XP, S=S, H=H, partials=partials)


@_on_tensor_device
def bn_finalize(partials, which: int, count: int, gamma, beta, eps: float, momentum: float,
                running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor]):
    """Batch statistics -> (scale, shift, mean, rstd); updates running stats in place."""
    lib = _lib.load()
    rows, nq, d = partials.shape
    require_cuda(partials, gamma, beta, running_mean, running_var)
    out = torch.empty(4, d, device=partials.device, dtype=torch.float32)
    _lib.check(lib.alignn_b200_bn_finalize(ptr(partials), rows, nq * d, which, count, d, ptr(gamma), ptr(beta), eps,
                                            momentum, ptr(running_mean), ptr(running_var), ptr(out[0]), ptr(out[1]),
                                            ptr(out[2]), ptr(out[3]), stream_ptr()), "alignn_b200_bn_finalize")
    return out[0], out[1], out[2], out[3]


@_on_tensor_device
def affine_silu_residual(R, res, scale, shift):
    lib = _lib.load()
    n, d = R.shape
    _check_d(d)
    require_cuda(R, res, scale, shift)
    out = torch.empty_like(R)
    with _span("affine_silu_residual", 4 * n * d * (2 + (res is not None))):
        _lib.check(lib.alignn_b200_affine_silu_residual(ptr(R), ptr(res), ptr(scale), ptr(shift), ptr(out), n, d,
                                                         stream_ptr()), "alignn_b200_affine_silu_residual")
    return out


@_on_tensor_device
def colsum(a: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """Deterministic fp64-accumulated column sum of a 2-D fp32 tensor (used on partial buffers)."""
    lib = _lib.load()
    require_cuda(a)
    rows, cols = a.shape
    out = torch.empty(cols, device=a.device, dtype=torch.float32)
    _lib.check(lib.alignn_b200_colsum(ptr(a), rows, cols, cols, alpha, ptr(out), stream_ptr()), "alignn_b200_colsum")
    return out


class BNLink:
    """Travels (as a Python attribute) on the OUTPUT tensor of a train-mode `BatchNorm1d -> SiLU (+ residual)` stage to
    the Linear that consumes it.  That Linear's data-gradient GEMM produces exactly the gradient the BatchNorm backward
    has to reduce first (c1 = mean gu, c2 = mean gu*xhat), so its epilogue accumulates those sums on the way out
    (`gemm_gather(bn_aux=...)`) and leaves them here; the producing stage's backward then skips its own reduction pass
    over the two [rows, d] tensors.  The producer only trusts the sums if the gradient it receives is the very tensor
    that GEMM wrote (`grad_ptr`): if autograd summed several consumers it falls back to `bn_backward_reduce`."""
    __slots__ = ("rows", "scale", "shift", "mean", "rstd", "n", "partials", "grad_ptr")

    def __init__(self, rows, scale, shift, mean, rstd, n):
        self.rows, self.scale, self.shift, self.mean, self.rstd, self.n = rows, scale, shift, mean, rstd, int(n)
        self.partials = None
        self.grad_ptr = None

    def usable_for(self, x: torch.Tensor) -> bool:
        return self.rows is not None and tuple(self.rows.shape) == tuple(x.shape) and self.rows.device == x.device

    def take(self, g_out: torch.Tensor):
        """(c1, c2) if the consumer's GEMM left sums for exactly this gradient tensor, else None."""
        part, self.partials = self.partials, None
        if part is None or g_out is None or g_out.data_ptr() != self.grad_ptr:
            return None
        return bn_backward_finish(part, self.n, self.rstd)


# Measured on B200 (batch 64, d = 256): the data-gradient GEMM with the two extra reductions in its epilogue takes 535 us
# instead of 225 us -- the epilogue of the tensor-core kernels is already the saturated agent (shared-memory / L1 pipe,
# DESIGN.md section 4) -- while the reduction pass it replaces costs 121 us.  Off by default; kept for A/B runs.
USE_BN_LINKS = os.environ.get("ALIGNN_B200_BN_LINKS", "0") == "1"


def bn_backward_finish(partials: torch.Tensor, n: int, rstd: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(c1, c2) from the partial sums a `gemm_gather(..., bn_aux=...)` epilogue produced:
    c1 = sum(gu) / n,  c2 = rstd * sum(gu * (m - mean)) / n."""
    rows, two, d = partials.shape
    c = colsum(partials.view(rows, 2 * d), 1.0 / n)
    return c[:d], c[d:] * rstd


@_on_tensor_device
def bn_backward_reduce(R, g_out, scale, shift, mean, rstd) -> Tuple[torch.Tensor, torch.Tensor]:
    """c1 = mean(gu), c2 = mean(gu*xhat) per channel (BatchNorm train-mode backward, pass 1)."""
    lib = _lib.load()
    n, d = R.shape
    _check_d(d)
    require_cuda(R, g_out, scale, shift, mean, rstd)
    rows = partial_rows(n, d)
    partials = torch.empty(rows, 2 * d, device=R.device, dtype=torch.float32)
    with _span("bn_backward_reduce", 8 * n * d):
        _lib.check(lib.alignn_b200_bn_backward_reduce(ptr(R), ptr(g_out), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), n, d,
                                                       ptr(partials), rows, stream_ptr()), "alignn_b200_bn_backward_reduce")
    c = colsum(partials, 1.0 / n)
    return c[:d], c[d:]


@_on_tensor_device
def egc_backward(ix: EdgeIndex, P, M, XP, S, H, gx_out, gy_out, n, e, *, reduce: bool = True, norm_nodes: int, norm_edges: int,
                 gate_eps: float = 1e-6, ln_eps: float = 1e-5):
    """n / e: dicts with keys w, b, mean, rstd, c1, c2 (entries may be None).

    Returns GM [Ne,d], GP [Nn,4d], vec_dst [6,d], vec_src [2,d] (column sums of the partials); with reduce=False the
    per-block partial rows [rows, 6d] and [rows, 2d] themselves."""
    lib = _lib.load()
    Nn, d = XP.shape
    Ne = M.shape[0]
    _check_d(d)
    require_cuda(P, M, XP, S, H, gx_out, gy_out, ix.src, ix.dst, ix.in_ptr, ix.in_eid, ix.out_ptr, ix.out_eid,
                 *[t for dct in (n, e) for t in dct.values()])
    dev = XP.device
    new = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
    GM, GP, GSh = new(Ne, d), new(Nn, 4 * d), new(Nn, d)
    rows = partial_rows(Nn, d)
    part, part_src = new(rows, 6 * d), new(rows, 2 * d)
    g = lambda dct, k: ptr(dct.get(k))  # noqa: E731
    a = _lib.EgcBwdArgs(
        struct_size=C.sizeof(_lib.EgcBwdArgs), Nn=Nn, Ne=Ne, d=d, norm_nodes=norm_nodes, norm_edges=norm_edges,
        gate_eps=gate_eps, ln_eps=ln_eps, P=ptr(P), M=ptr(M), XP=ptr(XP), S=ptr(S), H=ptr(H),
        src=ptr(ix.src), dst=ptr(ix.dst), in_ptr=ptr(ix.in_ptr), in_eid=None if ix.dst_sorted else ptr(ix.in_eid),
        out_ptr=ptr(ix.out_ptr), out_eid=ptr(ix.out_eid),
        n_w=g(n, "w"), n_b=g(n, "b"), n_mean=g(n, "mean"), n_rstd=g(n, "rstd"),
        e_w=g(e, "w"), e_b=g(e, "b"), e_mean=g(e, "mean"), e_rstd=g(e, "rstd"),
        n_c1=g(n, "c1"), n_c2=g(n, "c2"), e_c1=g(e, "c1"), e_c2=g(e, "c2"),
        gx_out=ptr(gx_out), gy_out=ptr(gy_out), GM=ptr(GM), GP=ptr(GP), GSh=ptr(GSh),
        partials=ptr(part), partials_src=ptr(part_src), stream=stream_ptr())
    # destination-keyed pass reads M, gy_out, node rows and writes GM, GP; source-keyed pass reads GM, M: SURVEY 8d
    nb = 4 * d * (Ne * (2 + (gy_out is not None)) + 9 * Nn) + 12 * Ne + 4 * d * 2 * Ne
    with _span("egc_backward(dst+src)", nb):
        _lib.check(lib.alignn_b200_egc_backward(C.byref(a)), "alignn_b200_egc_backward")
    if not reduce:              # the caller sums the per-block partial rows itself (WgradQueue: one batched launch per backward)
        return GM, GP, part, part_src
    return GM, GP, colsum(part).view(6, d), colsum(part_src).view(2, d)


@_on_tensor_device
def gather_segment_sum(ix: EdgeIndex, Bh, sigma):
    """Sh[v] = sum_{e->v} Bh[src e] * sigma[e];  S[v] = sum_{e->v} sigma[e]  (alignn.py:105-108)."""
    lib = _lib.load()
    Nn, d = Bh.shape
    Ne = sigma.shape[0]
    _check_d(d)
    require_cuda(Bh, sigma, ix.src, ix.in_ptr, ix.in_eid)
    Sh, S = torch.empty_like(Bh), torch.empty_like(Bh)
    _lib.check(lib.alignn_b200_gather_segment_sum(ptr(Bh), ptr(sigma), ptr(ix.src), ptr(ix.in_ptr),
                                                   None if ix.dst_sorted else ptr(ix.in_eid), Nn, Ne, d, ptr(Sh), ptr(S),
                                                   stream_ptr()), "alignn_b200_gather_segment_sum")
    return Sh, S


@_on_tensor_device
def pair_force_scatter(pair_forces: torch.Tensor, ix: EdgeIndex, add_reverse: bool = True) -> torch.Tensor:
    """forces[v] = sum over in-edges of pair_forces - (add_reverse ? sum over out-edges : 0): DGL's
    update_all(copy_e, sum) on g and on dgl.reverse(g) (alignn_atomwise.py:547-563) in one deterministic kernel."""
    lib = _lib.load()
    pf = pair_forces.contiguous()
    require_cuda(pf, ix.in_ptr, ix.in_eid, ix.out_ptr, ix.out_eid)
    Nn = ix.in_ptr.numel() - 1
    out = torch.empty(Nn, 3, device=pf.device, dtype=torch.float32)
    _lib.check(lib.alignn_b200_pair_force_scatter(ptr(pf), ptr(ix.in_ptr), None if ix.dst_sorted else ptr(ix.in_eid), ptr(ix.out_ptr),
                                                  ptr(ix.out_eid), Nn, int(add_reverse), ptr(out), stream_ptr()),
               "alignn_b200_pair_force_scatter")
    return out


@_on_tensor_device
def virial_stress(r: torch.Tensor, pair_forces: torch.Tensor, edge_offsets64: torch.Tensor, node_offsets64: torch.Tensor,
                  V: torch.Tensor, multiplier: float = 1.0) -> torch.Tensor:
    """stress[b] = multiplier * -160.21766208 * (r_b^T F_b) / V[first atom of b], one block per crystal
    (alignn_atomwise.py:610-635)."""
    lib = _lib.load()
    r, pf, V = r.contiguous(), pair_forces.contiguous(), V.contiguous().to(torch.float32)
    B = edge_offsets64.numel() - 1
    out = torch.empty(B, 3, 3, device=r.device, dtype=torch.float32)
    _lib.check(lib.alignn_b200_virial_stress(ptr(r), ptr(pf), edge_offsets64.data_ptr(), node_offsets64.data_ptr(), ptr(V), B,
                                             float(multiplier), ptr(out), stream_ptr()), "alignn_b200_virial_stress")
    return out


class _SegmentMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gptr):
        lib = _lib.load()
        require_cuda(x, gptr)
        B, d = gptr.numel() - 1, x.shape[1]
        out = torch.empty(B, d, device=x.device, dtype=torch.float32)
        _lib.check(lib.alignn_b200_segment_mean(ptr(x), ptr(gptr), B, d, ptr(out), stream_ptr()), "alignn_b200_segment_mean")
        ctx.save_for_backward(gptr)
        ctx.n = x.shape[0]
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        lib = _lib.load()
        (gptr,) = ctx.saved_tensors
        g_out = g_out.contiguous()
        B, d = g_out.shape
        gx = torch.empty(ctx.n, d, device=g_out.device, dtype=torch.float32)
        _lib.check(lib.alignn_b200_segment_mean_backward(ptr(g_out), ptr(gptr), B, d, ptr(gx), stream_ptr()),
                   "alignn_b200_segment_mean_backward")
        return gx, None


@_on_tensor_device
def segment_mean(x: torch.Tensor, graph_ptr: torch.Tensor) -> torch.Tensor:
    """Per-graph mean over node rows (dgl.nn.AvgPooling, alignn.py:325)."""
    return _SegmentMean.apply(x.contiguous(), graph_ptr)


def segment_mean_any_order(x: torch.Tensor, graph_ptr: torch.Tensor, second_order: bool) -> torch.Tensor:
    """`segment_mean`, or (force training, which differentiates through the backward) the same pooling from torch operators."""
    if not second_order:
        return segment_mean(x, graph_ptr)
    counts = (graph_ptr[1:] - graph_ptr[:-1]).long()
    gid = torch.repeat_interleave(torch.arange(counts.numel(), device=x.device), counts, output_size=x.shape[0])
    sums = torch.zeros(counts.numel(), x.shape[1], device=x.device, dtype=x.dtype).index_add(0, gid, x)
    return sums / counts.clamp_min(1).to(x.dtype).unsqueeze(1)


# ---- tensor-core Linear (tcgen05, bf16x3) --------------------------------------------------------
class WeightImage:
    """bf16 hi/lo image of a weight matrix W[N,K] (or of W^T when transpose=True) in UMMA core-matrix
    order; valid until the weight changes (rebuilt every step in training)."""

    __slots__ = ("buf", "N", "K")

    def __init__(self, W: torch.Tensor, transpose: bool = False):
        lib = _lib.load()
        require_cuda(W)
        if W.dim() != 2:
            raise RuntimeError("WeightImage needs a 2-D weight")
        self.N, self.K = (W.shape[1], W.shape[0]) if transpose else (W.shape[0], W.shape[1])
        nbytes = int(lib.alignn_b200_gemm_weight_image_bytes(self.N, self.K))
        if nbytes == 0:
            raise RuntimeError(f"alignn_b200 GEMM: unsupported weight shape N={self.N}, K={self.K} (need multiples of 32)")
        self.buf = torch.empty(nbytes, device=W.device, dtype=torch.uint8)
        _lib.check(lib.alignn_b200_gemm_prepare_weights(ptr(W), self.N, self.K, W.stride(0), int(transpose), ptr_any(self.buf),
                                                        stream_ptr()), "alignn_b200_gemm_prepare_weights")


class _Img:
    """An operand image owned by an ImageTable (same attributes as WeightImage)."""
    __slots__ = ("buf", "N", "K")

    def __init__(self, N: int, K: int, device):
        nbytes = int(_lib.load().alignn_b200_gemm_weight_image_bytes(N, K))
        if nbytes == 0:
            raise RuntimeError(f"alignn_b200 GEMM: unsupported weight shape N={N}, K={K} (need multiples of 32)")
        self.N, self.K = N, K
        self.buf = torch.zeros(nbytes, device=device, dtype=torch.uint8)     # zero: K padding stays zero forever


class ImageTable:
    """bf16 hi/lo operand images (and stacked / folded bias vectors) of a group of Linear layers, rebuilt by ONE
    table-driven launch (`alignn_b200_gemm_prepare_table`) when -- and only when -- a source tensor changed
    (storage pointer or autograd version).  Tables register their blocks once; a parent table (the model) absorbs
    the tables of its layers so that a training step refreshes every image of the model in a single launch.

    Inside a CUDA-graph capture the refresh of a table with trainable sources is always recorded: a captured training
    step must rebuild its images on every replay, whatever the version counters said at capture time."""

    def __init__(self, device=None):
        self.device = device
        self.images = {}
        self.vectors = {}
        self._blocks = []      # (image name, source tensor, transpose, n_off, k_off)
        self._biases = []      # (vector name, offset, n, a, b)
        self._children = []
        self._parent = None
        self._key = None
        self._dev = None       # (entries tensor, n_entries, max_units, bias tensor, n_bias, pointer key)

    # -- construction ---------------------------------------------------------------------------
    def add_image(self, name: str, N: int, K: int, blocks, device):
        """blocks: [(W, transpose, n_off, k_off)]; W is a 2-D fp32 tensor (a Linear weight)."""
        self.images[name] = _Img(N, K, device)
        for W, tr, n_off, k_off in blocks:
            if k_off % 8:
                raise RuntimeError("ImageTable: k_off must be a multiple of 8")
            self._blocks.append((name, W, bool(tr), int(n_off), int(k_off)))
        self._key = self._dev = None

    def add_vector(self, name: str, n: int, parts, device):
        """parts: [(offset, a, b_or_None)]: vector[offset : offset + len(a)] = a (+ b)."""
        self.vectors[name] = torch.zeros(n, device=device, dtype=torch.float32)
        for off, a, b in parts:
            self._biases.append((name, int(off), int(a.numel()), a, b))
        self._key = self._dev = None

    def absorb(self, child: "ImageTable"):
        self._children.append(child)
        child._parent = weakref.ref(self)
        self._key = self._dev = None

    # -- refresh ----------------------------------------------------------------------------------
    def _all(self):
        out = [self]
        for c in self._children:
            out.extend(c._all())
        return out

    def _sources(self):
        for t in self._all():
            for _, W, _, _, _ in t._blocks:
                yield W
            for _, _, _, a, b in t._biases:
                yield a
                if b is not None:
                    yield b

    def _build_device_table(self):
        ents, bias = [], []
        max_units = 1
        for t in self._all():
            for name, W, tr, n_off, k_off in t._blocks:
                if W.dim() != 2 or W.stride(1) != 1 or not W.is_cuda or W.dtype != torch.float32:
                    raise RuntimeError("ImageTable: sources must be 2-D fp32 CUDA tensors with unit column stride")
                img = t.images[name]
                rows, cols = W.shape
                n_img, k_img = (cols, rows) if tr else (rows, cols)
                if n_off + n_img > img.N or k_off + (k_img + 7) // 8 * 8 > img.K:
                    raise RuntimeError(f"ImageTable: block does not fit image {name}")
                ents.append(_lib.ImageEntry(W=W.data_ptr(), ldw=W.stride(0), rows=rows, cols=cols, transpose=int(tr), n_off=n_off,
                                            k_off=k_off, N=img.N, K=img.K, image=img.buf.data_ptr()))
                max_units = max(max_units, n_img * ((k_img + 7) // 8))
            for name, off, n, a, b in t._biases:
                dst = t.vectors[name]
                bias.append(_lib.BiasEntry(a=a.data_ptr(), b=None if b is None else b.data_ptr(),
                                           dst=dst.data_ptr() + 4 * off, n=n))
        dev = next(self._sources()).device

        def pack(items, cls):
            if not items:
                return None
            arr = (cls * len(items))(*items)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory()
            self._host_keepalive.append(host)       # a captured copy reads this buffer again on every replay
            return host.to(dev, non_blocking=True)
        self._host_keepalive = []
        self._dev = (pack(ents, _lib.ImageEntry), len(ents), max_units, pack(bias, _lib.BiasEntry), len(bias),
                     tuple(s.data_ptr() for s in self._sources()))

    def refresh(self):
        """Rebuild the images if any source changed.  Returns True if a launch was issued."""
        srcs = list(self._sources())
        if not srcs:
            return False
        key = tuple((s.data_ptr(), s._version) for s in srcs)
        capturing = torch.cuda.is_current_stream_capturing() and any(s.requires_grad for s in srcs)
        if capturing and self._parent is not None and self._parent() is not None:
            capturing = False                       # the model-level table records the refresh of all its layers
        if key == self._key and not capturing:
            return False
        ptr_key = tuple(k[0] for k in key)
        if self._dev is None or self._dev[5] != ptr_key:
            self._build_device_table()
        ents, n_e, max_units, bias, n_b, _ = self._dev
        with torch.cuda.device(srcs[0].device):
            _lib.check(_lib.load().alignn_b200_gemm_prepare_table(ptr_any(ents), n_e, max_units, ptr_any(bias), n_b, stream_ptr()),
                       "alignn_b200_gemm_prepare_table")
        for t in self._all():
            t._key = tuple((s.data_ptr(), s._version) for s in t._sources())
        return True


def ptr_any(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


@_on_tensor_device
def gemm_nt(A: torch.Tensor, w: WeightImage, bias: Optional[torch.Tensor] = None,
            residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = A[M,K] @ W^T (+ bias) (+ residual) on tcgen05.  A may be a column slice (row stride >= K)."""
    lib = _lib.load()
    if A.dim() != 2 or A.stride(1) != 1 or A.shape[1] != w.K:
        raise RuntimeError(f"gemm_nt: A must be [M,{w.K}] with unit column stride, got {tuple(A.shape)}/{A.stride()}")
    for t in (A, bias, residual, out):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError("gemm_nt needs fp32 CUDA tensors")
    M = A.shape[0]
    if out is None:
        out = torch.empty(M, w.N, device=A.device, dtype=torch.float32)
    ldr = residual.stride(0) if residual is not None else 0
    if residual is not None and (residual.stride(1) != 1 or residual.shape != (M, w.N)):
        raise RuntimeError("gemm_nt: bad residual layout")
    with _span(f"gemm_nt<{min(w.N, 256)}>", 4 * M * (w.K + w.N * (1 + (residual is not None)))):
        _lib.check(lib.alignn_b200_gemm_nt(ptr_any(A), A.stride(0), ptr_any(w.buf), M, w.N, w.K, ptr_any(bias), ptr_any(residual),
                                            ldr, ptr_any(out), out.stride(0), stream_ptr()), "alignn_b200_gemm_nt")
    return out


@_on_tensor_device
def gemm_gather(A: torch.Tensor, w: WeightImage, bias: Optional[torch.Tensor] = None, *,
                add0: Optional[torch.Tensor] = None, idx0: Optional[torch.Tensor] = None,
                add1: Optional[torch.Tensor] = None, idx1: Optional[torch.Tensor] = None,
                stats: bool = False, out: Optional[torch.Tensor] = None, bn_aux=None):
    """out[r] = A[r] @ W^T (+ bias) (+ add0[idx0[r]]) (+ add1[idx1[r]]) on tcgen05, A streamed by TMA tiles.

    add0 / add1 are 2-D fp32 views with unit column stride and w.N columns (column slices of a wider matrix are fine);
    idx None = identity (a residual).  stats=True also returns the per-CTA partial column sums [rows, 2, N] of out and
    out^2 (alignn.py:123 batch statistics; feed `bn_finalize(partials, 0, M, ...)`).
    bn_aux = (m_rows [M,N], scale, shift, mean): `out` is the gradient w.r.t. the output of a train-mode
    BatchNorm1d + SiLU whose pre-norm rows are m_rows; returns (out, partials [rows,2,N]) with the partial sums of
    gu = out * silu'(m*scale+shift) and gu * (m - mean) (feed `bn_backward_finish`)."""
    lib = _lib.load()
    if A.dim() != 2 or A.stride(1) != 1 or A.shape[1] != w.K:
        raise RuntimeError(f"gemm_gather: A must be [M,{w.K}] with unit column stride, got {tuple(A.shape)}/{A.stride()}")
    M = A.shape[0]
    for t in (A, bias, add0, add1, out):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError("gemm_gather needs fp32 CUDA tensors")
    for t, ix in ((add0, idx0), (add1, idx1)):
        if t is None:
            if ix is not None:
                raise RuntimeError("gemm_gather: index without addend")
            continue
        if t.dim() != 2 or t.stride(1) != 1 or t.shape[1] != w.N:
            raise RuntimeError("gemm_gather: addend must be [rows, N] with unit column stride")
        if ix is None and t.shape[0] != M:
            raise RuntimeError("gemm_gather: identity-indexed addend must have M rows")
        if ix is not None and (ix.dtype != torch.int32 or not ix.is_cuda or not ix.is_contiguous() or ix.numel() != M):
            raise RuntimeError("gemm_gather: index must be a contiguous int32 CUDA tensor with M entries")
    if out is None:
        out = torch.empty(M, w.N, device=A.device, dtype=torch.float32)
    part = None
    bn_ptrs = (None, None, None)
    if bn_aux is not None:
        if add1 is not None or stats:
            raise RuntimeError("gemm_gather: bn_aux excludes add1 / stats")
        add1, sc, sh, mu = bn_aux
        if add1.shape != (M, w.N) or add1.stride(1) != 1:
            raise RuntimeError("gemm_gather: bn_aux rows must be [M, N]")
        require_cuda(sc, sh, mu)
        bn_ptrs = (ptr(sc), ptr(sh), ptr(mu))
        stats = True
    if stats:
        rows = int(lib.alignn_b200_gemm_gather_stat_rows(M, w.N))
        part = torch.empty(max(rows, 1), 2, w.N, device=A.device, dtype=torch.float32)
    a = _lib.GemmGatherArgs(
        struct_size=C.sizeof(_lib.GemmGatherArgs), M=M, N=w.N, K=w.K, A=ptr_any(A), lda=A.stride(0),
        w_image=ptr_any(w.buf), bias=ptr_any(bias),
        add0=ptr_any(add0), ld0=add0.stride(0) if add0 is not None else 0, idx0=ptr_any(idx0),
        add1=ptr_any(add1), ld1=add1.stride(0) if add1 is not None else 0, idx1=ptr_any(idx1),
        C=ptr_any(out), ldc=out.stride(0), stats=ptr_any(part), bn_scale=bn_ptrs[0], bn_shift=bn_ptrs[1], bn_mean=bn_ptrs[2],
        stream=stream_ptr())
    nb = 4 * M * (w.K + w.N) + (8 * M if idx0 is not None else 0) + (4 * M * w.N if bn_aux is not None else 0) + (4 * M * w.N if (add0 is not None and idx0 is None) else 0)
    kind = "+gather" if idx0 is not None else ("+residual" if add0 is not None else "")
    if bn_aux is not None:
        kind += "+bn_bwd"
    with _span(f"gemm_gather<{min(w.N, 256)}>" + kind + ("+stats" if stats else ""), nb):
        _lib.check(lib.alignn_b200_gemm_gather(C.byref(a)), "alignn_b200_gemm_gather")
    return (out, part) if stats else out


def wgrad_supported(DA: int, DB: int) -> bool:
    return int(_lib.load().alignn_b200_wgrad_workspace_bytes(32, DA, DB, 1)) > 0


@_on_tensor_device
def wgrad(A: torch.Tensor, B: torch.Tensor, groups: int = 1) -> torch.Tensor:
    """out[g*DA + o, i] = sum_r A[r, g*DA + o] * B[r, i]  (dL/dW of a Linear: A = output grads, B = inputs)."""
    lib = _lib.load()
    require_cuda(A, B)
    K, DB = B.shape
    if A.dim() != 2 or A.shape[0] != K or A.shape[1] % groups:
        raise RuntimeError(f"wgrad: A must be [{K}, groups*DA], got {tuple(A.shape)}")
    DA = A.shape[1] // groups
    nbytes = int(lib.alignn_b200_wgrad_workspace_bytes(K, DA, DB, groups))
    if nbytes == 0 and K >= 0:
        raise RuntimeError(f"alignn_b200 wgrad: unsupported shape DA={DA}, DB={DB}")
    out = torch.empty(groups * DA, DB, device=A.device, dtype=torch.float32)
    ws = torch.empty(max(nbytes, 16), device=A.device, dtype=torch.uint8)
    with _span(f"wgrad<{DA},{DB}>", 4 * K * (groups * DA + DB)):
        _lib.check(lib.alignn_b200_wgrad(ptr(A), A.stride(0), ptr(B), B.stride(0), K, DA, DB, groups, ptr(out), DB, ptr_any(ws),
                                          nbytes, stream_ptr()), "alignn_b200_wgrad")
    return out


WGRAD_BATCH_MAX = 64      # problems per launch (kMaxProblems in csrc/wgrad_tc.cu)


@_on_tensor_device
def wgrad_batch(problems) -> None:
    """ONE launch for many square weight gradients: `problems` is a list of (A [K, >= d] view, B [K, >= d] view, out [d, d]
    view); out[o, i] = sum_r A[r, o] * B[r, i].  Views may be column slices of wider matrices (row stride = stride(0))."""
    lib = _lib.load()
    if not problems:
        return
    d = problems[0][2].shape[0]
    for i0 in range(0, len(problems), WGRAD_BATCH_MAX):
        chunk = problems[i0:i0 + WGRAD_BATCH_MAX]
        arr = (_lib.WgradProblem * len(chunk))()
        nbytes_in = 0
        for q, (A, B, out) in zip(arr, chunk):
            if not (A.is_cuda and B.is_cuda and out.is_cuda and A.dtype == B.dtype == out.dtype == torch.float32):
                raise RuntimeError("alignn_b200 kernels need fp32 CUDA tensors")
            if out.shape != (d, d) or A.shape[1] != d or B.shape[1] != d or A.shape[0] != B.shape[0] or A.stride(1) != 1 \
                    or B.stride(1) != 1 or out.stride(1) != 1:
                raise RuntimeError("wgrad_batch: every problem is A [K, d], B [K, d] -> out [d, d] with unit column stride")
            q.A, q.lda, q.B, q.ldb, q.K = ptr(A), A.stride(0), ptr(B), B.stride(0), A.shape[0]
            q.out, q.ld_out = ptr(out), out.stride(0)
            nbytes_in += 8 * A.shape[0] * d
        nbytes = int(lib.alignn_b200_wgrad_batch_workspace_bytes(arr, len(chunk), d))
        if nbytes == 0:
            raise RuntimeError(f"alignn_b200 wgrad_batch: unsupported batch (d={d}, n={len(chunk)})")
        # the non-cooperative fallback runs problem by problem and needs the single-problem workspace
        nbytes = max(nbytes, max(int(lib.alignn_b200_wgrad_workspace_bytes(A.shape[0], d, d, 1)) for A, _, _ in chunk))
        ws = torch.empty(nbytes, device=chunk[0][0].device, dtype=torch.uint8)
        with _span(f"wgrad_batch<{d}>", nbytes_in):
            _lib.check(lib.alignn_b200_wgrad_batch(arr, len(chunk), d, ptr_any(ws), nbytes, stream_ptr()), "alignn_b200_wgrad_batch")


class WgradQueue:
    """Deferred weight gradients.  While a queue is installed (`WgradQueue.current`), the conv Functions do not launch their
    weight-gradient GEMMs during backward; they register (A, B, destination) here and return None for those weights, and
    `flush()` computes all of them with one `wgrad_batch` launch, writing straight into the destinations -- slices of
    the flat gradient buffer of `alignn_b200.dp.FlatGradAllReducer`, which installs the queue.  A weight is only deferred
    if the queue knows a destination for it (`dest`: parameter data_ptr -> [d, d] view) and it has not been queued already
    in this backward (a layer applied twice falls back to the immediate path and autograd's accumulation)."""
    current: Optional["WgradQueue"] = None

    def __init__(self):
        self.dest = {}            # weight.data_ptr() -> destination view [d, d]
        self.vec_dest = {}        # bias / norm parameter data_ptr() -> destination view [d]
        self.items = []           # (A, B, out)
        self.vec_items = []       # (partial rows [rows, n*d], column offset, d, out)
        self._seen = set()

    def wants(self, *weights) -> bool:
        keys = [w.data_ptr() for w in weights]
        return all(k in self.dest and k not in self._seen for k in keys) and len(set(keys)) == len(keys)

    def wants_vecs(self, *params) -> bool:
        keys = [p.data_ptr() for p in params]
        return all(k in self.vec_dest and k not in self._seen for k in keys) and len(set(keys)) == len(keys)

    def add(self, A: torch.Tensor, B: torch.Tensor, weight: torch.Tensor) -> None:
        k = weight.data_ptr()
        self._seen.add(k)
        self.items.append((A, B, self.dest[k]))

    def add_vec(self, partials: torch.Tensor, block: int, d: int, param: torch.Tensor) -> None:
        """param.grad = column sums of partials[:, block*d:(block+1)*d] (per-block partial rows of egc_backward)."""
        k = param.data_ptr()
        self._seen.add(k)
        self.vec_items.append((partials, block * d, d, self.vec_dest[k]))

    def deferred_ptrs(self):
        return self._seen

    def flush(self) -> None:
        items, self.items = self.items, []
        vecs, self.vec_items = self.vec_items, []
        self._seen = set()
        wgrad_batch(items)
        colsum_batch(vecs)


@_on_tensor_device
def colsum_batch(problems) -> None:
    """ONE launch for many column sums: `problems` = list of (partials [rows, >= off + d], off, d, out [d])."""
    if not problems:
        return
    lib = _lib.load()
    arr = (_lib.ColsumProblem * len(problems))()
    for q, (part, off, d, out) in zip(arr, problems):
        if not (part.is_cuda and out.is_cuda and part.dtype == out.dtype == torch.float32) or part.stride(1) != 1 or out.numel() != d:
            raise RuntimeError("colsum_batch: fp32 CUDA partial rows with unit column stride and a [d] output")
        q.a, q.rows, q.stride, q.cols, q.alpha, q.out = part.data_ptr() + 4 * off, part.shape[0], part.stride(0), d, 1.0, ptr(out)
    _lib.check(lib.alignn_b200_colsum_batch(arr, len(problems), stream_ptr()), "alignn_b200_colsum_batch")


@_on_tensor_device
def colsum_rows(a: torch.Tensor) -> torch.Tensor:
    """Column sums of a tall contiguous [n, d] matrix, deterministic two-stage reduction."""
    lib = _lib.load()
    require_cuda(a)
    n, d = a.shape
    _check_d(d)
    rows = partial_rows(n, d)
    part = torch.empty(rows, d, device=a.device, dtype=torch.float32)
    _lib.check(lib.alignn_b200_colsum_partials(ptr(a), n, d, ptr(part), rows, stream_ptr()), "alignn_b200_colsum_partials")
    return colsum(part)


def linear_table(lin) -> ImageTable:
    """Images of one nn.Linear (weight zero-padded along K to a multiple of 32) and of its transpose, cached on the module."""
    dev = lin.weight.device
    tbl = getattr(lin, "_alignn_b200_images", None)
    if tbl is not None and tbl.device == dev:
        return tbl
    k_pad = (lin.in_features + 31) // 32 * 32
    tbl = ImageTable(dev)
    tbl.add_image("w", lin.out_features, k_pad, [(lin.weight, False, 0, 0)], dev)
    tbl.add_image("wT", k_pad, lin.out_features, [(lin.weight, True, 0, 0)], dev)
    object.__setattr__(lin, "_alignn_b200_images", tbl)
    return tbl


class input_grads_only:
    """Context manager for a backward pass whose PARAMETER gradients are thrown away -- `torch.autograd.grad(energy, r)`
    for forces (alignn/models/alignn_atomwise.py:530-539 outside force training): the library's autograd Functions then
    skip their weight-gradient GEMMs and bias / norm-parameter reductions and return None for them.  (autograd's own
    `needs_input_grad` cannot tell: it is fixed at forward time from `requires_grad`.)  A plain process-wide flag, because
    the autograd engine runs CUDA nodes on its own thread."""
    active = False

    def __enter__(self):
        self._prev = input_grads_only.active
        input_grads_only.active = True

    def __exit__(self, *exc):
        input_grads_only.active = self._prev


def _pad_cols(x: torch.Tensor, k_pad: int) -> torch.Tensor:
    x = x.contiguous()
    return x if x.shape[1] == k_pad else torch.nn.functional.pad(x, (0, k_pad - x.shape[1]))


class _TCLinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 bf16x3 GEMMs (forward, data gradient, weight gradient)."""

    @staticmethod
    def forward(ctx, x, weight, bias, tbl):
        x = _pad_cols(x, tbl.images["w"].K)
        ctx.save_for_backward(x)
        ctx.tbl = tbl
        ctx.k_in = weight.shape[1]
        return gemm_gather(x, tbl.images["w"], bias.contiguous())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        (x,) = ctx.saved_tensors
        go = go.contiguous()
        gx = None
        if ctx.needs_input_grad[0]:
            gx = gemm_gather(go, ctx.tbl.images["wT"])[:, :ctx.k_in]
        if input_grads_only.active:
            return gx, None, None, None
        gw = wgrad(go, x, 1)[:, :ctx.k_in]
        gb = colsum_rows(go)
        return gx, gw, gb, None


def tc_linear_supported(in_features: int, out_features: int) -> bool:
    k_pad = (in_features + 31) // 32 * 32
    return out_features in _lib.SUPPORTED_D and wgrad_supported(out_features, k_pad)


@_on_tensor_device
def tc_linear(x: torch.Tensor, lin) -> torch.Tensor:
    """nn.Linear forward/backward on the tensor-core kernels; the input width is zero-padded to a multiple of 32."""
    tbl = linear_table(lin)
    tbl.refresh()
    return _TCLinearFn.apply(x, lin.weight, lin.bias, tbl)


# ---- Linear -> BatchNorm1d(train) -> SiLU (embedding MLP layers) ------------------------------------
class _MLPBNTrainFn(torch.autograd.Function):
    """One embedding layer in train mode, entirely on library kernels: tensor-core Linear, two-stage batch
    statistics (fp64 finalize + running-stat update), fused normalise+SiLU, and the matching backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, bn, tbl):
        x = _pad_cols(x, tbl.images["w"].K)
        # Linear + per-channel batch statistics in one pass (column sums leave through the GEMM epilogue)
        R, part = gemm_gather(x, tbl.images["w"], bias.contiguous(), stats=True)
        n, d = R.shape
        track = bn.track_running_stats and bn.running_mean is not None
        scale, shift, mean, rstd = bn_finalize(part, 0, n, gamma.contiguous(), beta.contiguous(), bn.eps, float(bn.momentum),
                                               bn.running_mean if track else None, bn.running_var if track else None)
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        out = affine_silu_residual(R, None, scale, shift)
        ctx.save_for_backward(x, R, scale, shift, mean, rstd)
        ctx.link = BNLink(R, scale, shift, mean, rstd, n) if USE_BN_LINKS else None
        ctx.tbl = tbl
        ctx.k_in = weight.shape[1]
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        lib = _lib.load()
        x, R, scale, shift, mean, rstd = ctx.saved_tensors
        go = go.contiguous()
        n, d = R.shape
        got = ctx.link.take(go) if ctx.link is not None else None     # sums left by the consumer's data-gradient GEMM
        c1, c2 = got if got is not None else bn_backward_reduce(R, go, scale, shift, mean, rstd)
        gR = torch.empty_like(R)
        with _span("bn_backward_apply", 12 * n * d):
            _lib.check(lib.alignn_b200_bn_backward_apply(ptr(R), ptr(go), ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
                                                         ptr(c1.contiguous()), ptr(c2.contiguous()), n, d, ptr(gR), stream_ptr()),
                       "alignn_b200_bn_backward_apply")
        gx = None
        if ctx.needs_input_grad[0]:
            gx = gemm_gather(gR, ctx.tbl.images["wT"])[:, :ctx.k_in]
        if input_grads_only.active:
            return gx, None, None, None, None, None, None
        gw = wgrad(gR, x, 1)[:, :ctx.k_in]
        # a bias that feeds a train-mode BatchNorm has an identically zero gradient (sum_rows gR == 0)
        return gx, gw, torch.zeros_like(c1), c2 * n, c1 * n, None, None


# ---- Linear -> LayerNorm -> SiLU (embedding layers of the LayerNorm model, alignn_atomwise.py:249-268) ---------------
class _MLPLNFn(torch.autograd.Function):
    """Tensor-core Linear, then ONE row kernel for LayerNorm + SiLU (forward) and one for their backward (which also
    leaves the per-block sums for d gamma / d beta); replaces 5 library passes over the [T, d] activations."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, eps, tbl):
        lib = _lib.load()
        x = _pad_cols(x, tbl.images["w"].K)
        h = gemm_gather(x, tbl.images["w"], bias.contiguous())
        n, d = h.shape
        gamma, beta = gamma.contiguous(), beta.contiguous()
        out = torch.empty_like(h)
        rowstat = torch.empty(n, 2, device=h.device, dtype=torch.float32)
        with _span("ln_silu_forward", 8 * n * d):
            _lib.check(lib.alignn_b200_ln_silu_forward(ptr(h), ptr(gamma), ptr(beta), float(eps), n, d, ptr(out), ptr(rowstat),
                                                        stream_ptr()), "alignn_b200_ln_silu_forward")
        ctx.save_for_backward(x, h, rowstat, gamma, beta)
        ctx.tbl = tbl
        ctx.k_in = weight.shape[1]
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        lib = _lib.load()
        x, h, rowstat, gamma, beta = ctx.saved_tensors
        go = go.contiguous()
        n, d = h.shape
        gh = torch.empty_like(h)
        rows = partial_rows(n, d)
        part = torch.empty(rows, 2 * d, device=h.device, dtype=torch.float32)
        with _span("ln_silu_backward", 12 * n * d):
            _lib.check(lib.alignn_b200_ln_silu_backward(ptr(h), ptr(go), ptr(rowstat), ptr(gamma), ptr(beta), n, d, ptr(gh),
                                                         ptr(part), rows, stream_ptr()), "alignn_b200_ln_silu_backward")
        gx = None
        if ctx.needs_input_grad[0]:
            gx = gemm_gather(gh, ctx.tbl.images["wT"])[:, :ctx.k_in]
        if input_grads_only.active:
            return gx, None, None, None, None, None, None
        gwb = colsum(part)
        gw = gb = None
        if ctx.needs_input_grad[1]:
            gw = wgrad(gh, x, 1)[:, :ctx.k_in]
        if ctx.needs_input_grad[2]:
            gb = colsum_rows(gh)
        return gx, gw, gb, gwb[:d], gwb[d:], None, None


@_on_tensor_device
def mlp_ln(x, lin, ln):
    tbl = linear_table(lin)
    tbl.refresh()
    return _MLPLNFn.apply(x, lin.weight, lin.bias, ln.weight, ln.bias, ln.eps, tbl)


@_on_tensor_device
def mlp_bn_train(x, lin, bn):
    tbl = linear_table(lin)
    tbl.refresh()
    out = _MLPBNTrainFn.apply(x, lin.weight, lin.bias, bn.weight, bn.bias, bn, tbl)
    if out.grad_fn is not None and getattr(out.grad_fn, "link", None) is not None:
        out._alignn_b200_bn_link = out.grad_fn.link
    return out
