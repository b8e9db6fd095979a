"""Developer tool: CUDA-event timings of individual library calls at the headline shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from alignn_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


big = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # 256 MB scratch to flush L2 between groups


for (M, N, K) in [(276480, 256, 256), (23040, 1024, 256), (23040, 256, 1024), (1920, 1024, 256), (1920, 256, 1024),
                  (23040, 256, 256), (276480, 64, 64)]:
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    img = ops.WeightImage(W)
    t = timeit(lambda: ops.gemm_nt(A, img, b))
    t_ref = timeit(lambda: torch.addmm(b, A, W.t()))
    flops = 2.0 * M * N * K
    byts = 4.0 * (M * K + M * N)
    print(f"gemm_nt M={M} N={N} K={K}: {t:8.1f} us  ({flops / t / 1e6:7.1f} TFLOP/s fp32-equiv, {byts / t / 1e3:7.1f} GB/s)   "
          f"torch fp32 addmm {t_ref:8.1f} us")
for (K, D, G) in [(276480, 256, 1), (23040, 256, 4), (23040, 256, 1), (1920, 256, 4)]:
    A = torch.randn(K, G * D, device=dev)
    B = torch.randn(K, D, device=dev)
    t = timeit(lambda: ops.wgrad(A, B, G))
    t_ref = timeit(lambda: A.t() @ B)
    print(f"wgrad K={K} D={D} groups={G}: {t:8.1f} us ({4.0 * K * D * (G + 1) / t / 1e3:7.1f} GB/s)   torch fp32 {t_ref:8.1f} us")
# residual epilogue (data-gradient GEMMs: gy = gy_out + GM W)
for (M, N, K) in [(276480, 256, 256), (23040, 256, 1024)]:
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    R = torch.randn(M, N, device=dev)
    img = ops.WeightImage(W)
    t = timeit(lambda: ops.gemm_nt(A, img, None, R))
    print(f"gemm_nt+residual M={M} N={N} K={K}: {t:8.1f} us ({4.0 * (M * K + 2 * M * N) / t / 1e3:7.1f} GB/s)")
