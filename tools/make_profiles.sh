#!/usr/bin/env bash
# Reproduce the evidence under profiles/ for round NN on a B200 box (run from the repo root, e.g. through
# `gpurun -- bash tools/make_profiles.sh 02`).  Nothing printed under ncu is ever used as a bench value.
set -uo pipefail
R=${1:-02}
OUT=gpurun_out
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.build()"
# headline bench lines (CUDA events, graphs, clocks sampled during the timed region)
python bench.py                                    | tail -1 > "$OUT/r${R}_bench_n1.json"
python bench.py --norm layernorm --no-cpu-baseline | tail -1 > "$OUT/r${R}_bench_n1_layernorm.json"
python bench.py --no-graph --no-cpu-baseline       | tail -1 > "$OUT/r${R}_bench_n1_eager.json"
# per-kernel ranking of one step (CUPTI) and the secondary BASELINE configs
python tools/profile_step.py  2>/dev/null | grep -v Warn | head -40 > "$OUT/r${R}_step_kernel_table_torch_profiler.txt"
python tools/bench_configs.py | tail -1 > "$OUT/r${R}_configs_2_and_5.json"
python tools/bench_config4.py | tail -1 > "$OUT/r${R}_config4_alignn_ff.json"
python tools/bench_gemm_gather.py | tail -1 > "$OUT/r${R}_gemm_gather_ab.json"
./tools/microbench > "$OUT/r${R}_microbench_tma_l2.jsonl" 2>&1 || true
# ncu: full sections for every hot kernel at the L(g) shape (one training conv: forward + backward, second repetition)
ncu --set full --clock-control none --import-source on \
    -k regex:"egc_|wgrad_bf16x3|gemm_gather|affine_silu|bn_backward_reduce" -s 12 -c 12 \
    -o "$OUT/r${R}_final_bn" -f python tools/run_kernels.py conv_bn 2 > "$OUT/ncu_full.log" 2>&1
# ... and the kernels that do not occur in a single conv: batched weight gradients, LayerNorm + SiLU rows, flat AdamW
ncu --set full --clock-control none --import-source on \
    -k regex:"wgrad_batch|ln_silu|adamw_flat" -s 3 -c 4 \
    -o "$OUT/r${R}_final_extras" -f python tools/run_kernels.py extras 2 > "$OUT/ncu_full_extras.log" 2>&1
python tools/ncu_summary.py "$OUT/r${R}_final_bn.ncu-rep" "$OUT/r${R}_final_extras.ncu-rep" > "$OUT/r${R}_ncu_full_summary.md"
python tools/ncu_traffic.py "$OUT/r${R}_final_bn.ncu-rep" "$OUT/r${R}_final_extras.ncu-rep" > "$OUT/r${R}_ncu_traffic.json"
python tools/trace_gemm.py plain > "$OUT/r${R}_gemm_trace.txt" 2>&1
python tools/trace_gemm.py gate >> "$OUT/r${R}_gemm_trace.txt" 2>&1
python tools/bench_egc_ring.py | tail -1 > "$OUT/r${R}_egc_ring_ab.json"
# launch list of one training step (eager launches; graph replays launch the same kernels)
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file "$OUT/r${R}_launches_all.csv" python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline \
    > "$OUT/ncu_launches.log" 2>&1
python tools/last_step_launches.py "$OUT/r${R}_launches_all.csv" "$OUT/r${R}_launches_bench_step.csv"
python tools/launch_table.py "$OUT/r${R}_launches_bench_step.csv" > "$OUT/r${R}_launch_table.txt"
# sanitizers on a subset of the parity tests
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm.py -q -x \
    -k "jvasp or edge_cases or edgeless or 128-256-256 or 127-64-64 or 384-32-64 or 130-32-4 or 3-256-1 or 777-64-32" > "$OUT/r${R}_compute_sanitizer_memcheck.log" 2>&1
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm.py -q -x \
    -k "jvasp or edgeless or 127-64-64 or 384-32-64 or 130-32-4" > "$OUT/r${R}_compute_sanitizer_racecheck.log" 2>&1
echo "copy $OUT/r${R}_* (except *.ncu-rep and *_launches_all.csv) into profiles/ and commit"
