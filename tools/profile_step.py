"""Developer tool: per-kernel GPU time of a few training steps via torch.profiler (CUPTI).
Not a benchmark -- numbers under a profiler are never reported; use it to rank kernels."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from alignn_b200 import dp, synthetic  # noqa: E402
from alignn_b200.alignn import ALIGNN, ALIGNNConfig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--norm", default="batchnorm")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--infer", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(123)
if args.norm == "layernorm":
    from alignn_b200 import alignn_atomwise as AW

    class Model(ALIGNN):
        _mlp, _alignn_conv, _gcn_conv = AW.MLPLayer, AW.ALIGNNConv, AW.EdgeGatedGraphConv
    model = Model(ALIGNNConfig(name="alignn"))
else:
    model = ALIGNN(ALIGNNConfig(name="alignn"))
model.to(dev).train(not args.infer)
g, lg, lat, tgt = [t.to(dev) for t in synthetic.make_batch(64, 30, 12, seed=123)]
# the step bench.py times: flat gradient buffer, weight gradients deferred into one batched launch, one-launch AdamW
reducer = dp.FlatGradAllReducer(model.parameters())
opt = None


def step():
    if args.infer:
        with torch.no_grad():
            return model((g, lg, lat))
    reducer.zero_grad()
    loss = (model((g, lg, lat)) - tgt).abs().mean()
    with reducer.deferring():
        loss.backward()
    reducer.all_reduce()
    if opt is not None:
        opt.step()


step()
if not args.infer:
    opt = dp.FlatAdamW(reducer, lr=1e-3)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / args.steps, e.count / args.steps) for e in prof.key_averages() if e.device_time_total > 0
        and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total GPU kernel time per step: {tot / 1e3:.3f} ms")
for k, t, c in rows[:40]:
    print(f"{t / tot * 100:6.2f}% {t:10.1f} us/step  n/step={c:6.1f}  {k[:110]}")
