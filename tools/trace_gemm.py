"""Developer tool: per-role event timeline of CTA 0 of the gather GEMM (SM clocks), to find which agent paces the tile.
roles: 0 = A box issued, 1 = W chunk issued, 2 = converter sees staged box, 3 = converter published planes,
4 = MMA thread issues a K chunk, 5 = epilogue: accumulator ready / accumulator released (alternating)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from alignn_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
d, E, T = 256, 23040, 276480
g = torch.Generator().manual_seed(1)
y = torch.randn(T, d, generator=g).to(dev)
W = (torch.randn(d, d, generator=g) / 16).to(dev)
P = torch.randn(E, 4 * d, generator=g).to(dev)
dst = torch.arange(E).repeat_interleave(12).to(torch.int32).to(dev)
src = ((dst.cpu().long() // 360) * 360 + torch.randint(0, 360, (T,), generator=g)).to(torch.int32).to(dev)
img = ops.WeightImage(W)
out = torch.empty(T, d, device=dev)


def run():
    if mode == "plain":
        ops.gemm_gather(y, img, None, out=out)
    else:
        ops.gemm_gather(y, img, None, add0=P[:, 0:d], idx0=src, add1=P[:, 2 * d:3 * d], idx1=dst, stats=True, out=out)


lib = _lib.load()
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = torch.zeros(6, 512, dtype=torch.int64, device=dev)
lib._handle  # noqa: B018
fn = lib.alignn_b200_debug_gemm_trace
fn.argtypes = [C.c_void_p]
fn.restype = None
fn(buf.data_ptr())
run()
torch.cuda.synchronize()
fn(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
names = ["A box issue", "W chunk issue", "conv: box seen", "conv: planes out", "MMA chunk issue", "epi: acc ready/released"]
for r in range(6):
    row = [int(v) - t0 for v in t[r].tolist() if v > 0][:40]
    print(f"{names[r]:26s}", " ".join(f"{v:6d}" for v in row))
# steady state: cycles per tile from the MMA chunk issues (8 chunks per tile)
m = [int(v) - t0 for v in t[4].tolist() if v > 0]
per_tile = [m[i + 8] - m[i] for i in range(0, len(m) - 8, 8)]
print("cycles per tile (MMA issue to MMA issue):", per_tile[:16])
e = [int(v) - t0 for v in t[5].tolist() if v > 0]
# per tile: ready, then per chunk (after tcgen05.ld, after staging), then released = 2 + 16 stamps
NC = 4
per = 2 + 2 * NC
for ti in range(3):
    seg = e[ti * per:(ti + 1) * per]
    if len(seg) == per:
        print(f"tile {ti}: ready {seg[0]}; per chunk (ld wait, staging, rest):",
              [(seg[1 + 2 * c] - (seg[2 * c] if c else seg[0]), seg[2 + 2 * c] - seg[1 + 2 * c],
                (seg[3 + 2 * c] if c < NC - 1 else seg[per - 1]) - seg[2 + 2 * c]) for c in range(NC)])
