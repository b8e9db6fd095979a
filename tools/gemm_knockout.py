"""Developer tool: time the gate GEMM with one pipeline agent disabled at a time."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from alignn_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
M, N, K = 276480, 256, 256
A = torch.randn(M, K, device=dev)
W = torch.randn(N, K, device=dev) / 16
b = torch.randn(N, device=dev)
img = ops.WeightImage(W)


def t(flags):
    lib.alignn_b200_debug_gemm_flags(ctypes.c_int(flags))
    for _ in range(2):
        ops.gemm_nt(A, img, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.gemm_nt(A, img, b)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 * 1e3


for name, f in [("full (L2 tile prefetch)", 0), ("full, no L2 prefetch", 2048), ("no C", 4), ("no C, no L2 prefetch", 4 + 2048)]:
    print(f"{name:40s} {t(f):8.1f} us")
for Mx in (276480 // 4, 276480 // 16, 128 * 148):
    A = torch.randn(Mx, K, device=dev)
    print("M =", Mx, "tiles/SM =", Mx / 128 / 148)
    for name, f in [("full", 0), ("nothing at all", 63)]:
        print(f"   {name:37s} {t(f):8.1f} us")
lib.alignn_b200_debug_gemm_flags(ctypes.c_int(0))
