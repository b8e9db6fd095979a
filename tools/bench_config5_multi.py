"""BASELINE config 5 on N GPUs of one node: the gather / segment-sum primitive (DGL update_all(u_mul_e,sum) +
update_all(copy_e,sum), alignn.py:105-108) on 1e5 / 1e6 / 1e7 edges PER GPU, d = 256.

The primitive never crosses a crystal, so the edge partitions are independent: every rank owns
one partition, there is no data-path collective ("weak" scaling).  Timing: barrier + synchronize, CUDA events on every
rank, MAX over ranks; value = total algorithmic bytes of all ranks / that time.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_config5_multi.py          (or plain `python tools/bench_config5_multi.py` for N = 1)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from alignn_b200 import ops  # noqa: E402
from alignn_b200.graph import Graph  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0


def timed(fn, reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


out = {}
for ne in (100_000, 1_000_000, 10_000_000):
    # the index as BASELINE config 5 defines it (synthetic.make_segment_sweep: fan-in 12, uniform random sources); the
    # feature values are drawn on the device (10 GB of host random numbers per rank at 1e7 edges would dominate the run)
    gen = torch.Generator().manual_seed(123 + rank)
    nn_ = max(1, ne // 12)
    dst = torch.repeat_interleave(torch.arange(nn_), 12)
    src = torch.randint(0, nn_, (nn_ * 12,), generator=gen)
    g = Graph(src, dst, nn_)
    gd = g.to(dev)
    dgen = torch.Generator(device=dev).manual_seed(7 + rank)
    sigma = torch.rand(nn_ * 12, 256, device=dev, generator=dgen)
    bh = torch.randn(nn_, 256, device=dev, generator=dgen)

    def run():
        return ops.gather_segment_sum(gd.index, bh, sigma)
    for _ in range(5):
        run()
    ms = timed(run, 20 if ne < 10_000_000 else 10)
    nbytes = 1288.0 * g.num_edges() * world
    out[str(ne)] = {"edges_per_gpu": g.num_edges(), "us": ms * 1e3, "GBps_all_gpus": nbytes / (ms * 1e-3) / 1e9,
                    "frac_of_measured_hbm_per_gpu": nbytes / world / (ms * 1e-3) / 1e9 / peak}
    del g, gd, bh, sigma
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps({"config5_gather_segment_sum": out, "n_gpus": world, "scaling": "weak", "d": 256,
                      "bytes_per_edge": 1288, "timing": "CUDA events, max over ranks"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
