#!/usr/bin/env bash
# Validation and A/B timing of the experimental fully fused kernels (alignn_b200/csrc/staged).
#   gpurun --timeout 1500 -- 'bash tools/staged_first_run.sh'
# Every step runs under its own `timeout` (the mbarrier waits of the fused kernel trap after 2^26 spins instead of
# hanging, but a wedged context must not eat the whole box); logs land in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export ALIGNN_B200_STAGED=1
python tools/build_staged.py > gpurun_out/staged_build.log 2>&1 || { echo "staged build failed"; tail -5 gpurun_out/staged_build.log; exit 1; }
# 2. the fused forward, smallest configuration first
timeout 300 python -m pytest tests/test_staged.py -m gpu -q -k "fused and not backward" > gpurun_out/staged_fused_tests.log 2>&1
rc=$?
echo "fused forward: exit $rc"; tail -15 gpurun_out/staged_fused_tests.log
# 2b. the fused backward (independent of the forward kernel's result)
timeout 300 python -m pytest tests/test_staged.py -m gpu -q -k "fused_backward" > gpurun_out/staged_bwd_tests.log 2>&1
rcb=$?
echo "fused backward: exit $rcb"; tail -15 gpurun_out/staged_bwd_tests.log
if [ $rcb -eq 0 ]; then
  timeout 300 python tools/bench_fused_bwd.py >> gpurun_out/staged_bench_fused.jsonl 2>> gpurun_out/staged_bench_fused.err
fi
# 3. timing only if it is correct
if [ $rc -eq 0 ]; then
  for mode in stats affine layer; do
    for groups in 1 2; do
      timeout 300 python tools/bench_fused.py --mode $mode --groups $groups >> gpurun_out/staged_bench_fused.jsonl 2>> gpurun_out/staged_bench_fused.err
    done
  done
  cat gpurun_out/staged_bench_fused.jsonl
  # 4. one ncu --set full capture of the fused kernel per variant (first matching launch only; the script is tiny, so
  #    the replay passes cost seconds, not minutes)
  for groups in 1 2; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:egc_forward_fused -c 1 \
      -o gpurun_out/staged_fused_g${groups} -f python tools/bench_fused.py --mode stats --groups $groups --iters 2 --graphs lg \
      > gpurun_out/staged_ncu_g${groups}.log 2>&1
    echo "ncu groups=$groups: exit $?"
  done
fi
