"""Developer tool: run the hot kernels at the headline L(g) shapes a few times (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from alignn_b200 import ops, synthetic  # noqa: E402
from alignn_b200.alignn import EdgeGatedGraphConv  # noqa: E402
from alignn_b200.alignn_atomwise import EdgeGatedGraphConv as ConvLN  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
g, lg, lat, tgt = synthetic.make_batch(64, 30, 12, seed=123)
lg = lg.to(dev)
E, T, d = lg.num_nodes(), lg.num_edges(), 256
m = torch.randn(E, d, device=dev)
z = torch.randn(T, d, device=dev)

if which in ("gemm", "all"):
    W = torch.randn(d, d, device=dev) / 16
    b = torch.randn(d, device=dev)
    img = ops.WeightImage(W)
    for _ in range(reps):
        out = ops.gemm_nt(z, img, b)
if which in ("wgrad", "all"):
    for _ in range(reps):
        out = ops.wgrad(z, z, 1)
if which in ("conv_bn", "conv_ln", "all"):
    for cls, name in ((EdgeGatedGraphConv, "conv_bn"), (ConvLN, "conv_ln")):
        if which not in (name, "all"):
            continue
        conv = cls(d, d).to(dev).train()
        mi = m.clone().requires_grad_(True)
        zi = z.clone().requires_grad_(True)
        for _ in range(reps):
            xo, yo = conv(lg, mi, zi)
            (xo.sum() + yo.sum()).backward()
torch.cuda.synchronize()
print("done", which)
