#!/usr/bin/env bash
# Multi-GPU evidence for round NN on ONE box with 8 GPUs (`gpurun --gpus 8 -- bash tools/multi_gpu_evidence.sh 02`):
# BASELINE config 5 (gather / segment-sum primitive) at 1 / 2 / 4 / 8 GPUs and the headline bench at 2 / 4 / 8 GPUs, the
# latter also with the gradient all-reduce captured inside the CUDA graph.  Every run is bounded by `timeout`.
set -uo pipefail
R=${1:-02}
OUT=gpurun_out
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -c "import __graft_entry__ as g; g.build()"
timeout 300 python tools/bench_config5_multi.py | tail -1 > "$OUT/r${R}_config5_n1.json"
port=29531
for n in 2 4 8; do
  port=$((port + 1))
  timeout 300 $TR --nproc-per-node $n --master-port $port tools/bench_config5_multi.py 2> "$OUT/config5_n$n.err" | tail -1 > "$OUT/r${R}_config5_n$n.json"
done
for n in ${BENCH_NS:-8}; do
  port=$((port + 1))
  timeout 400 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 20 --warmup 3 2> "$OUT/bench_n$n.err" | tail -1 > "$OUT/r${R}_bench_n$n.json"
done
port=$((port + 1))
[ "${IN_GRAPH:-0}" = "1" ] && ALIGNN_B200_NCCL_IN_GRAPH=1 timeout 400 $TR --nproc-per-node 8 --master-port $port bench.py --gpus 8 --steps 20 --warmup 3 2> "$OUT/bench_n8_ig.err" | tail -1 > "$OUT/r${R}_bench_n8_allreduce_in_graph.json"
python - "$OUT" "$R" <<'PY'
import json, sys
out, r = sys.argv[1], sys.argv[2]
for n in (1, 2, 4, 8):
    try:
        j = json.load(open(f"{out}/r{r}_config5_n{n}.json"))
        print("config5 n=%d" % n, {k: round(v["GBps_all_gpus"]) for k, v in j["config5_gather_segment_sum"].items()})
    except Exception as e:
        print("config5 n=%d: no result (%s)" % (n, e))
for tag in ("n2", "n4", "n8", "n8_allreduce_in_graph"):
    try:
        j = json.load(open(f"{out}/r{r}_bench_{tag}.json"))
        print("bench", tag, round(j["value"], 1), "graphs/s", round(j["ms_per_step"], 3), "ms/step; e2e", round(j["e2e"]["value"], 1))
    except Exception as e:
        print("bench %s: no result (%s)" % (tag, e))
PY
