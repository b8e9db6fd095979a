"""Developer tool: A/B timing of the TMA-fed gather GEMM (csrc/gemm_fused_tc.cu) against the register-fed gemm_nt
(+ the separate passes it replaces) at the headline L(g) shapes.  CUDA events, L2 flushed between iterations."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from alignn_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def set_pair(on):
    from alignn_b200 import _lib
    import ctypes as C
    f = _lib.load().alignn_b200_debug_gemm_pair
    f.argtypes, f.restype = [C.c_int], None
    f(int(on))


def main():
    d = 256
    E, T = 23040, 276480
    g = torch.Generator().manual_seed(1)
    res = {}
    y = torch.randn(T, d, generator=g).to(dev)
    W = (torch.randn(d, d, generator=g) / 16).to(dev)
    b = torch.randn(d, generator=g).to(dev)
    P = torch.randn(E, 4 * d, generator=g).to(dev)
    dst = torch.arange(E).repeat_interleave(12).to(torch.int32).to(dev)
    src = ((dst.cpu().long() // 360) * 360 + torch.randint(0, 360, (T,), generator=g)).to(torch.int32).to(dev)
    R = torch.randn(T, d, generator=g).to(dev)
    img = ops.WeightImage(W)
    out = torch.empty(T, d, device=dev)
    hbm = 6385.8
    t = timeit(lambda: ops.gemm_nt(y, img, b, out=out))
    res["gemm_nt_lg"] = dict(us=round(t, 1), frac=round(4 * T * d * 2 / t / 1e3 / hbm, 3))
    for pair in (0, 1):
        set_pair(pair)
        tag = "pair" if pair else "1cta"
        t = timeit(lambda: ops.gemm_gather(y, img, b, out=out))
        res[f"gather_plain_lg_{tag}"] = dict(us=round(t, 1), frac=round(4 * T * d * 2 / t / 1e3 / hbm, 3))
        t = timeit(lambda: ops.gemm_gather(y, img, None, add0=P[:, 0:d], idx0=src, add1=P[:, 2 * d:3 * d], idx1=dst, stats=True, out=out))
        res[f"gather_gate_stats_lg_{tag}"] = dict(us=round(t, 1), frac=round(4 * T * d * 2 / t / 1e3 / hbm, 3))
        t = timeit(lambda: ops.gemm_gather(y, img, None, add0=R, out=out))
        res[f"gather_residual_lg_{tag}"] = dict(us=round(t, 1), frac=round(4 * T * d * 3 / t / 1e3 / hbm, 3))
        yg_ = torch.randn(E, d, generator=g).to(dev)
        t = timeit(lambda: ops.gemm_gather(yg_, img, b))
        res[f"gather_g_{tag}"] = dict(us=round(t, 1))
    t = timeit(lambda: ops.gemm_nt(y, img, None, R, out=out))
    res["gemm_nt_residual_lg"] = dict(us=round(t, 1), frac=round(4 * T * d * 3 / t / 1e3 / hbm, 3))
    t = timeit(lambda: ops.gemm_gather(y, img, None, add0=R, out=out))
    res["gather_residual_lg"] = dict(us=round(t, 1), frac=round(4 * T * d * 3 / t / 1e3 / hbm, 3))
    # node projections of an L(g) conv: [E,256] x [256,1024]
    x = torch.randn(E, d, generator=g).to(dev)
    Wc = (torch.randn(4 * d, d, generator=g) / 16).to(dev)
    bc = torch.randn(4 * d, generator=g).to(dev)
    imgc = ops.WeightImage(Wc)
    t = timeit(lambda: ops.gemm_nt(x, imgc, bc))
    res["gemm_nt_P"] = dict(us=round(t, 1))
    t = timeit(lambda: ops.gemm_gather(x, imgc, bc))
    res["gather_P"] = dict(us=round(t, 1))
    # g-graph sized
    yg = torch.randn(E, d, generator=g).to(dev)
    t = timeit(lambda: ops.gemm_nt(yg, img, b))
    res["gemm_nt_g"] = dict(us=round(t, 1))
    t = timeit(lambda: ops.gemm_gather(yg, img, b))
    res["gather_g"] = dict(us=round(t, 1))
    print(json.dumps(dict(bench="gemm_gather", **res)))


if __name__ == "__main__":
    main()
