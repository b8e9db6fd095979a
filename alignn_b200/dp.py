"""Data parallelism for the conv path: one process per GPU, graphs sharded across ranks, one
flat-buffer gradient all-reduce per step.

The reference wraps the model in `DistributedDataParallel(find_unused_parameters=True)`
(alignn/train.py:205-207) after `dist.init_process_group("nccl")` (train_alignn.py:33-38).  Batched
crystal graphs are block-diagonal, so forward and backward need no exchange at all (SURVEY.md
section 8e); the only collective is the gradient all-reduce: 4,026,753 fp32 = 16.1 MB per step for
the default model.  Here every gradient lives in ONE contiguous fp32 buffer (`param.grad` are views
into it), so the step issues a single all-reduce (NCCL over NVLink 5 / NVSwitch, in-switch NVLS
reduction when available) instead of DDP's bucket walk plus the unused-parameter graph traversal.

Parameters that never receive a gradient (the dead bn_edges affine pairs) keep
`grad is None`, exactly as under the reference's DDP, so optimizers skip them.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """Rendezvous from torchrun-style env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).

    Returns (rank, local_rank, world_size).  A single process (no env) is world_size 1 with no
    process group.
    """
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_range(num_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of `num_items` graphs for `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class FlatGradAllReducer:
    """Owns one flat fp32 gradient buffer for `params` and all-reduces (averages) it once per step.

    Usage per step:   reducer.zero_grad(); loss.backward(); reducer.all_reduce(); optimizer.step()
    (`all_reduce` = `gather()` into the flat buffer + the collective; with CUDA graphs call `gather()` inside the
    captured forward/backward graph and `reduce_flat()` eagerly between the graphs.)
    The set of parameters that receive gradients is discovered on the first backward.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        self.flat: Optional[torch.Tensor] = None
        self.views: List[torch.Tensor] = []
        self.active: List[torch.nn.Parameter] = []
        self._work = None
        self.queue = None            # ops.WgradQueue once `deferring()` has been used

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _build(self) -> None:
        self.active = [p for p in self.params if p.grad is not None]
        if not self.active:
            raise RuntimeError("no parameter received a gradient; call after the first backward()")
        dev, dt = self.active[0].device, self.active[0].dtype
        if any(p.dtype != dt or p.device != dev for p in self.active):
            raise RuntimeError("FlatGradAllReducer needs all gradients on one device with one dtype")
        n = sum(p.numel() for p in self.active)
        self.flat = torch.zeros(n, device=dev, dtype=dt)
        self.views = []
        off = 0
        for p in self.active:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self._refresh_queue()

    def _refresh_queue(self) -> None:
        """(Re)build the destination map of the deferred weight gradients: parameter storage -> its slice of the flat
        gradient buffer.  Called again by FlatAdamW after it re-homed the parameters (their data_ptr changed)."""
        if self.queue is None or self.flat is None or not self.flat.is_cuda:
            return
        self.queue.dest = {p.data_ptr(): v for p, v in zip(self.active, self.views)
                           if p.dim() == 2 and p.shape[0] == p.shape[1] and v.data_ptr() % 16 == 0}
        self.queue.vec_dest = {p.data_ptr(): v for p, v in zip(self.active, self.views) if p.dim() == 1}

    def deferring(self):
        """Context manager for `loss.backward()`: the conv layers queue their weight-gradient GEMMs instead of launching
        them one by one, and `gather()` runs the whole queue as ONE batched launch that writes straight into the flat
        buffer (alignn_b200.ops.WgradQueue / alignn_b200_wgrad_batch).  No effect before the flat buffer exists (first
        backward) or on CPU tensors."""
        from . import ops
        red = self

        class _Ctx:
            def __enter__(self):
                if red.queue is None:
                    red.queue = ops.WgradQueue()
                    red._refresh_queue()
                self.prev = ops.WgradQueue.current
                ops.WgradQueue.current = red.queue

            def __exit__(self, exc_type, *exc):
                ops.WgradQueue.current = self.prev
                if exc_type is not None:          # a failed backward: drop what it queued (its operands are garbage)
                    red.queue.items.clear()
                    red.queue.vec_items.clear()
                    red.queue.deferred_ptrs().clear()
        return _Ctx()

    def zero_grad(self) -> None:
        """Drop the gradients: autograd then WRITES fresh gradient tensors during backward instead of launching one
        `grad += new` kernel per parameter (~180 tiny launches per step for ALIGNN)."""
        for p in (self.params if self.flat is None else self.active):
            p.grad = None

    def gather(self) -> None:
        """After backward: copy every gradient into the flat buffer (one multi-tensor copy) and make `param.grad`
        alias its slice, so the collective and the optimizer both work on the flat buffer."""
        if self.flat is None:
            self._build()
        deferred = set()
        if self.queue is not None:
            deferred = set(self.queue.deferred_ptrs())
            self.queue.flush()                      # one batched launch; results land in their slices of `flat`
        views, grads = [], []
        for p, v in zip(self.active, self.views):
            if p.data_ptr() in deferred:
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                    v.add_(p.grad)                  # the weight was also used outside the queue in this backward
                continue
            if p.grad is None:
                raise RuntimeError("a parameter that used to receive a gradient did not get one this step")
            views.append(v)
            grads.append(p.grad)
        torch._foreach_copy_(views, grads)
        for p, v in zip(self.active, self.views):
            p.grad = v

    def nbytes(self) -> int:
        return 0 if self.flat is None else self.flat.numel() * self.flat.element_size()

    def all_reduce(self, async_op: bool = False):
        """Average gradients over ranks.  First call also builds the flat buffer."""
        self.gather()
        return self.reduce_flat(async_op)

    def reduce_flat(self, async_op: bool = False):
        """The collective alone (the flat buffer must already hold this step's gradients)."""
        w = self.world_size
        if w == 1:
            return None
        if self.flat.is_cuda and dist.get_backend(self.group) == "nccl":
            # NCCL averages inside the collective (in-switch reduction on NVSwitch when NVLS is up): no scaling kernel
            self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        else:
            self.flat.div_(w)          # gloo has no AVG: pre-scale, then SUM
            self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        return self._work

    def wait(self) -> None:
        if self._work is not None:
            self._work.wait()
            self._work = None


class FlatAdamW:
    """AdamW over ONE flat parameter: the parameters that receive gradients (`reducer.active`, discovered on the first
    backward) are re-homed as views into one contiguous fp32 buffer, their gradients already live in the reducer's
    flat buffer, so the whole update is ONE launch of the library's `alignn_b200_adamw_flat` kernel (row_kernels.cu;
    torch's fused AdamW needs ~10 launches of ~40 us for ALIGNN's 87 small tensors, and 2 x 60 us on one flat tensor
    because its multi-tensor chunking leaves most SMs idle).  Same arithmetic as `torch.optim.AdamW` per element
    (alignn/train.py:253-263 builds that optimizer); parameters without gradients are untouched, as with the per-parameter
    optimizer.  The step count lives on the device, so the launch can be captured in a CUDA graph and replayed
    (`capturable` is accepted for compatibility and ignored on CUDA).  CPU tensors (the gloo tests) use torch.optim.AdamW on
    the flat parameter.

    After an eager step the autograd version of every parameter is bumped, so caches keyed on it (operand images,
    alignn_b200.ops.ImageTable) see the update; inside a CUDA-graph capture those caches refresh unconditionally."""

    def __init__(self, reducer: FlatGradAllReducer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, capturable: bool = False):
        if reducer.flat is None:
            raise RuntimeError("FlatAdamW: run one backward() and reducer.gather() first (discovers the trainable set)")
        self.reducer = reducer
        self.params = list(reducer.active)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        flat = torch.empty_like(reducer.flat)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view_as(p)
                off += n
        self.flat_param = torch.nn.Parameter(flat)
        self.flat_param.grad = reducer.flat
        self.opt = None
        if flat.is_cuda:
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq = torch.zeros_like(flat)
            self.step_count = torch.zeros(1, dtype=torch.int64, device=flat.device)
            self._ticket = torch.zeros(1, dtype=torch.int32, device=flat.device)
        else:
            self.opt = torch.optim.AdamW([self.flat_param], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        reducer._refresh_queue()                    # parameters moved: deferred weight gradients are keyed by storage

    def step(self, zero_grad: bool = False):
        """One update; `zero_grad=True` also clears the flat gradient buffer in the same pass (CUDA)."""
        flat = self.flat_param.data
        if self.opt is None:
            from . import _lib
            from ._lib import ptr, stream_ptr
            with torch.cuda.device(flat.device):
                _lib.check(_lib.load().alignn_b200_adamw_flat(
                    ptr(flat), ptr(self.reducer.flat), ptr(self.exp_avg), ptr(self.exp_avg_sq), flat.numel(), self.lr,
                    self.betas[0], self.betas[1], self.eps, self.weight_decay, int(zero_grad), ptr(self.step_count),
                    ptr(self._ticket), stream_ptr()), "alignn_b200_adamw_flat")
        else:
            self.flat_param.grad = self.reducer.flat
            self.opt.step()
            if zero_grad:
                self.reducer.zero_grad()
        if not (flat.is_cuda and torch.cuda.is_current_stream_capturing()):
            for p in self.params:
                torch.autograd.graph.increment_version(p)

    def zero_grad(self):
        self.reducer.zero_grad()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers (what DDP does at wrap time)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)
