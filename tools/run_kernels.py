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
if which in ("extras", "all"):
    # the batched weight gradients of one training step (4 x T rows, 12 x E rows, 8 x 4 x N rows), the LayerNorm + SiLU row
    # kernels of an embedding layer at T rows, and the flat AdamW
    from alignn_b200 import dp
    N = g.num_nodes()
    gm, gp, x = torch.randn(T, d, device=dev), torch.randn(E, 4 * d, device=dev), torch.randn(N, d, device=dev)
    gpn = torch.randn(N, 4 * d, device=dev)
    probs = []
    for _ in range(4):
        probs.append((gm, z, torch.empty(d, d, device=dev)))
        probs += [(gp[:, j * d:(j + 1) * d], m, torch.empty(d, d, device=dev)) for j in range(4)]
    for _ in range(8):
        probs.append((m, m, torch.empty(d, d, device=dev)))
        probs += [(gpn[:, j * d:(j + 1) * d], x, torch.empty(d, d, device=dev)) for j in range(4)]
    layer = torch.nn.Sequential(torch.nn.Linear(64, d), torch.nn.LayerNorm(d), torch.nn.SiLU()).to(dev)
    from alignn_b200.alignn import mlp_forward
    h = torch.randn(T, 64, device=dev, requires_grad=True)
    flat = torch.nn.Parameter(torch.randn(4_100_000, device=dev))
    red = dp.FlatGradAllReducer([flat])
    flat.grad = torch.randn_like(flat)
    red.gather()
    opt = dp.FlatAdamW(red)
    for _ in range(reps):
        ops.wgrad_batch(probs)
        mlp_forward(layer, h).sum().backward()
        opt.step()
torch.cuda.synchronize()
print("done", which)
