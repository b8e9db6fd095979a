"""A/B timing of the staged one-kernel forward against the shipped gemm_nt + egc_forward pair (CUDA events, L2 flushed
between iterations) on the headline L(g) shape: batch 64, 30 atoms, k=12 -> 23 040 bonds, 276 480 bond pairs, d=256.

    python tools/bench_fused.py [--iters 20] [--d 256] [--batch 64]

Prints one JSON line: microseconds per call for each path and the algorithmic GB/s of the fused kernel
(read y + write M + node-sized tensors; P gathers count once)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--d", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--mode", choices=["stats", "affine", "layer"], default="stats")
    ap.add_argument("--groups", type=int, choices=[1, 2], default=1, help="epilogue groups of the fused kernel")
    ap.add_argument("--graphs", default="g,lg", help="which of the two graphs to time (comma separated: g, lg)")
    args = ap.parse_args()
    import staged_binding as SB
    from alignn_b200 import ops, synthetic
    lib = SB.load()
    dev = torch.device("cuda:0")
    g, lg, _, _ = synthetic.make_batch(batch_size=args.batch, atoms=30, k=12, seed=123)
    d = args.d
    res = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, gr in (("g", g), ("lg", lg)):
        if name not in args.graphs.split(","):
            continue
        grd = gr.to(dev)
        ix = grd.index
        Nn, Ne = gr.num_nodes(), gr.num_edges()
        gen = torch.Generator().manual_seed(1)
        x = torch.randn(Nn, d, generator=gen).to(dev)
        y = torch.randn(Ne, d, generator=gen).to(dev)
        Wcat = (torch.randn(4 * d, d, generator=gen) / d ** 0.5).to(dev)
        W_eg = (torch.randn(d, d, generator=gen) / d ** 0.5).to(dev)
        bcat, b_eg = torch.randn(4 * d, generator=gen).to(dev), torch.randn(d, generator=gen).to(dev)
        e_w, e_b = (torch.rand(d, generator=gen) + 0.5).to(dev), torch.randn(d, generator=gen).to(dev)
        P = ops.gemm_nt(x, ops.WeightImage(Wcat), bcat)
        img = ops.WeightImage(W_eg)
        n, tiles = SB.pack_tiles(lib, ix.in_ptr.cpu().numpy())
        tiles_d = torch.from_numpy(tiles).to(dev)
        norm = {"stats": ops.NORM_STATS, "affine": ops.NORM_AFFINE, "layer": ops.NORM_LAYER}[args.mode]
        train = args.mode != "affine"
        ones, zeros = torch.ones(d, device=dev), torch.zeros(d, device=dev)

        def fused(out=None):
            return SB.fused_forward(lib, ix, tiles_d, n, y, img, b_eg, P, norm, train, e_w, e_b, out=out, groups=args.groups)

        def shipped():
            G = ops.gemm_nt(y, img, b_eg)
            return ops.egc_forward(ix, x, y, G, P, ones, zeros, e_w, e_b, norm_nodes=ops.NORM_STATS if train else ops.NORM_AFFINE,
                                   norm_edges=norm, residual=True, save=train, need_edge_out=True)

        def timeit(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            return ts[len(ts) // 2]
        out = fused()
        t_f = timeit(lambda: fused(out))
        t_s = timeit(shipped)
        nbytes = 4 * d * (Ne * (1 + int(train) + (args.mode != "stats") * 2) + Nn * (4 + 1 + 2 * int(train))) + 12 * Ne
        res[name] = dict(Nn=Nn, Ne=Ne, tiles=n, fused_us=round(t_f, 1), shipped_us=round(t_s, 1),
                         fused_algorithmic_GBps=round(nbytes / t_f / 1e3, 1))
    print(json.dumps(dict(bench="fused_gate_forward", mode=args.mode, groups=args.groups, d=d, batch=args.batch, **res)))


if __name__ == "__main__":
    main()
