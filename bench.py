#!/usr/bin/env python
"""Benchmark of the ALIGNN edge-gated conv hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--norm batchnorm|layernorm]

A "step" is one forward + backward + optimizer update of ALIGNN (4 ALIGNN + 4 GCN layers, hidden
256, the `ALIGNN` class of alignn/models/alignn.py, L1 loss as in train.py:240) on one synthetic
JARVIS-DFT-shaped batch of 64 crystals per GPU (30 atoms, k=12: N=1920, E=23 040, T=276 480).
Weak scaling: every rank owns its own 64 graphs; the only collective is the gradient all-reduce.

Prints ONE JSON line (rank 0).  Keys follow the driver contract; extra keys:
  roofline      dominant kernel (fused edge kernel on L(g)): compulsory bytes / CUDA-event time
  step_hbm      whole-step compulsory bytes (10.04 GB per batch fwd+bwd) / step time
  cpu_baseline  the oracle (torch-CPU restatement of the reference DGL path) on this box's cores
  e2e           same metric with the batch starting in pinned HOST memory every step and the loss read back
`--impl reference` times that CPU oracle alone (the reference's own implementation needs DGL, which
cannot be installed offline).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "graphs/sec fwd+bwd"
UNIT = "graphs/s"
WORKLOAD = "configs[2]: ALIGNN training fwd+bwd+AdamW, batch=64 JARVIS-shaped graphs/GPU (30 atoms, k=12), 4+4 layers d=256"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--norm", default="batchnorm", choices=["batchnorm", "layernorm"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--atoms", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--cpu-sample-graphs", type=int, default=0, help="0 = calibrate (~4 s of CPU work per step)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# shared: byte model (SURVEY.md section 8d / BASELINE.md section 3)
# ---------------------------------------------------------------------------------------------
def conv_bytes_train(Nn, Ne, d):
    fwd = 4 * d * (2 * Nn + 2 * Ne) + 8 * Ne + 4 * d * Ne          # inference fwd + save m
    bwd = 4 * d * (5 * Nn + 4 * Ne) + 8 * Ne
    return fwd + bwd


def step_bytes(N, E, T, d, n_alignn, n_gcn):
    return n_alignn * (conv_bytes_train(N, E, d) + conv_bytes_train(E, T, d)) + n_gcn * conv_bytes_train(N, E, d)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            j = json.load(fh)
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------
def _oracle_setup(args, graphs):
    from alignn_b200 import synthetic
    from oracle import alignn_oracle as O
    g, lg, lat, tgt = synthetic.make_batch(batch_size=graphs, atoms=args.atoms, k=12, seed=123)

    def to_o(gr):
        s, d = gr.edges()
        og = O.OGraph(s.long(), d.long(), gr.num_nodes(), gr.batch_num_nodes(), gr.batch_num_edges())
        og.ndata.update(gr.ndata)
        og.edata.update(gr.edata)
        return og
    torch.manual_seed(123)
    model = O.ALIGNN(norm=args.norm)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    og, olg = to_o(g), to_o(lg)

    def step():
        opt.zero_grad(set_to_none=True)
        out = model((og, olg, lat))
        loss = (out - tgt).abs().mean()
        loss.backward()
        opt.step()
    return step


def cpu_calibrate(args):
    """Pick the host thread count the oracle runs fastest with (more threads is not faster for these
    gather/index_add-heavy ops on a 100+ core box) and a per-step sample size of about 4 s."""
    ncpu = os.cpu_count() or 1
    cands = sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)})
    step = _oracle_setup(args, 2)
    best = None
    for th in cands:
        torch.set_num_threads(th)
        step()
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (th, dt)
    threads, t2 = best
    graphs = int(max(1, min(args.batch, round(4.0 / (t2 / 2)))))
    return threads, graphs


def cpu_oracle_run(args, graphs, steps, warmup, threads):
    """graphs/s of the oracle (fwd+bwd+AdamW) on `graphs` crystals per step with `threads` host threads."""
    torch.set_num_threads(threads)
    step = _oracle_setup(args, graphs)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return graphs * len(times) / total, total / len(times) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads, graphs = cpu_calibrate(args)
    if args.cpu_sample_graphs > 0:
        graphs = args.cpu_sample_graphs
    gps, ms = cpu_oracle_run(args, graphs, args.steps, max(1, min(args.warmup, 3)), threads)
    cores = threads
    line = {
        "impl": "reference", "metric": METRIC, "value": gps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{graphs} graphs per step (bounded CPU sample of the 64-graph batch)",
                   "norm": args.norm},
        "cpu_baseline": {"value": gps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {graphs} graphs, torch-CPU restatement of the reference DGL path "
                                   f"(DGL is not installable offline); thread count calibrated over {os.cpu_count()} cores"},
        "e2e": {"value": gps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    from alignn_b200 import _lib, dp, ops, synthetic
    from alignn_b200.alignn import ALIGNN, ALIGNNConfig

    rank, local, world = dp.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()

    # ---- model ------------------------------------------------------------------------------
    torch.manual_seed(123)                                   # reference default seed, config.py:164
    cfg = ALIGNNConfig(name="alignn")
    if args.norm == "layernorm":
        from alignn_b200 import alignn_atomwise as AW

        class Model(ALIGNN):
            _mlp, _alignn_conv, _gcn_conv = AW.MLPLayer, AW.ALIGNNConv, AW.EdgeGatedGraphConv
        model = Model(cfg)
    else:
        model = ALIGNN(cfg)
    model.to(dev).train()
    dp.broadcast_parameters(model)
    use_graph = not args.no_graph
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True, capturable=use_graph)
    reducer = dp.FlatGradAllReducer(model.parameters())

    # ---- data: each rank owns its own batches (weak scaling); 4 distinct batches rotate ------
    nb = 4
    host = []
    for b in range(nb):
        g, lg, lat, tgt = synthetic.make_batch(batch_size=args.batch, atoms=args.atoms, k=12,
                                               seed=123 + 1000 * rank + b)
        host.append((g.pin_memory(), lg.pin_memory(), lat.pin_memory(), tgt.pin_memory()))
    N, E, T = host[0][0].num_nodes(), host[0][0].num_edges(), host[0][1].num_edges()
    resident = [(g.to(dev), lg.to(dev), lat.to(dev), tgt.to(dev)) for g, lg, lat, tgt in host]
    h2d_bytes = host[0][0].nbytes() + host[0][1].nbytes() + host[0][2].numel() * 4 + host[0][3].numel() * 4

    def step(batch):
        g, lg, lat, tgt = batch
        reducer.zero_grad()
        out = model((g, lg, lat))
        loss = (out - tgt).abs().mean()                      # nn.L1Loss, train.py:240
        loss.backward()
        reducer.all_reduce()
        opt.step()
        return loss

    def h2d(i):
        g, lg, lat, tgt = host[i % nb]
        return (g.to(dev, non_blocking=True), lg.to(dev, non_blocking=True), lat.to(dev, non_blocking=True),
                tgt.to(dev, non_blocking=True))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- warm-up (also builds the flat gradient buffer) ----------------------------------------
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(max(args.warmup, 3)):
            step(resident[i % nb])
    torch.cuda.current_stream().wait_stream(side)
    barrier()

    # ---- CUDA graphs: one per resident batch and one per host batch (H2D copies inside the graph) --------
    # The step is ~650 launches of which most are small (g-graph convs, norms, optimizer); replaying them as a
    # graph removes the host launch cost.  Shapes are static here; a real variable-size loader would bucket.
    # The gradient all-reduce stays OUTSIDE the graphs (an NCCL collective inside a captured graph hung on this
    # stack): per step = replay(zero_grad + forward + backward) -> eager flat all-reduce -> replay(optimizer).
    graphs_res, graphs_e2e, graph_opt, launches_per_step = [], [], None, None

    def fwd_bwd(batch):
        g, lg, lat, tgt = batch
        reducer.zero_grad()
        out = model((g, lg, lat))
        loss = (out - tgt).abs().mean()
        loss.backward()
        reducer.gather()                                      # gradients -> flat buffer (one multi-tensor copy)
        return loss

    if use_graph:
        pool = None
        for b in range(nb):
            gr = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(gr, pool=pool):
                loss_b = fwd_bwd(resident[b])
            launches_per_step = _lib.launch_count() - l0
            pool = pool or gr.pool()
            graphs_res.append((gr, loss_b))
        for b in range(nb):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, pool=pool):
                loss_b = fwd_bwd(h2d(b))
            graphs_e2e.append((gr, loss_b))
        graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_opt, pool=pool):
            opt.step()
        barrier()

    def run_resident(i):
        if use_graph:
            graphs_res[i % nb][0].replay()
            reducer.reduce_flat()
            graph_opt.replay()
        else:
            step(resident[i % nb])

    def run_e2e(i):
        if use_graph:
            gr, loss_b = graphs_e2e[i % nb]
            gr.replay()
            reducer.reduce_flat()
            graph_opt.replay()
            return loss_b.item()                              # D2H + sync, as train.py:300-305 does
        return step(h2d(i)).item()

    for i in range(2):
        run_resident(i)

    # ---- timed: resident inputs -------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms_total = timed(run_resident, args.steps)
    launches = (launches_per_step * args.steps) if use_graph else (_lib.launch_count() - l0)

    # ---- timed: end to end from pinned host memory, loss read back every step ----------------
    for i in range(2):
        run_e2e(i)
    ms_e2e = timed(run_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel: CUDA events around every L(g) launch of the fused edge kernel, eager replay of
    #      the same steps (events inside a graph replay carry no timestamps) ------------------------------
    for i in range(2):                                        # eager allocations settle on this stream
        step(resident[i % nb])
    ops.TIMER = ops.KernelTimer(min_edges=T // 2)
    ms_eager = timed(lambda i: step(resident[i % nb]), args.steps)
    ksum = ops.TIMER.summary()
    ops.TIMER = None

    lt = torch.tensor([launches], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(lt)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    graphs_per_step = args.batch * world
    value = graphs_per_step * args.steps / (ms_total / 1e3)
    e2e_value = graphs_per_step * args.steps / (ms_e2e / 1e3)
    peak, peak_src = peaks()
    d = cfg.hidden_features
    sbytes = step_bytes(N, E, T, d, cfg.alignn_layers, cfg.gcn_layers)
    ms_step = ms_total / args.steps
    k = ksum.get("egc_forward_kernel")
    roofline = None
    if k:
        ach = k["bytes"] / (k["avg_ms"] * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full` capture per round
        # (profiles/r01_ncu_full_summary.md: 383.6 + 302.6 MB in BatchNorm-train mode); unknown for other modes
        traffic = 686.2e6 if args.norm == "batchnorm" and (N, E, T) == (1920, 23040, 276480) else None
        roofline = {"bound": "hbm", "kernel": "egc_forward_kernel<256> on L(g) (Nn=E, Ne=T)", "achieved": ach,
                    "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                    "avg_launch_ms": k["avg_ms"], "launches_timed": k["launches"], "algorithmic_bytes_per_launch": k["bytes"]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "model": "ALIGNN 4+4 d=256 (" + args.norm + ", train mode)",
                   "global_batch": graphs_per_step, "per_gpu_batch": args.batch, "N": N, "E": E, "T": T,
                   "parallelism": f"dp{world}", "optimizer": "AdamW(fused)", "loss": "L1",
                   "cuda_graph": use_graph, "eager_ms_per_step": ms_eager / args.steps,
                   "l2": f"no explicit flush: per-step working set ~{sbytes / 1e9:.1f} GB >> 126 MB L2; 4 batches rotate"},
        "roofline": roofline,
        "step_hbm": {"algorithmic_bytes_per_step": sbytes, "achieved": sbytes / (ms_step * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": sbytes / (ms_step * 1e-3) / 1e9 / peak,
                     "note": "conv-stack compulsory bytes per batch (SURVEY 8d) / whole step time incl. embeddings, GEMMs, optimizer"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4},
        "gpu_launches": int(lt.item()),
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads, graphs = cpu_calibrate(args)
        if args.cpu_sample_graphs > 0:
            graphs = args.cpu_sample_graphs
        gps, ms = cpu_oracle_run(args, graphs, 3, 1, threads)
        line["cpu_baseline"] = {"value": gps, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"3 steps x {graphs} graphs (fwd+bwd+AdamW), torch-CPU restatement of the reference "
                                          f"DGL path; thread count calibrated over the box's {os.cpu_count()} cores"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
