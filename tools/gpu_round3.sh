#!/bin/bash
# Round-2 closing evidence run (one gpurun call, one GPU): full GPU test suite, smoke, the default bench line, warm per-kernel lists of
# the two teacher steps, and ncu captures of the teacher-step kernels (launch list + --set full).  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r2c}
timeout 1500 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err; echo "bench exit=$?"
timeout 300 python tools/prof_teacher_step.py bf16 2>&1 | grep -v -i warn > $OUT/${TAG}_teacher_warm_cfg2.txt
timeout 300 python tools/prof_teacher_step.py fp32 32 77 8 512 2>&1 | grep -v -i warn > $OUT/${TAG}_teacher_warm_cfg0.txt
for cfg in bf16 fp32; do
  CMD="python tools/run_teacher_step.py 3 $cfg"
  if [ $cfg = bf16 ]; then SKIP=18; CNT=9; else SKIP=30; CNT=15; fi
  timeout 300 $CMD > $OUT/${TAG}_ncu_plain_$cfg.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_teacher_$cfg.csv $CMD > $OUT/${TAG}_ncu_run_$cfg.log 2>&1
  echo "ncu launch list $cfg exit=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_attn_|k_split|k_sum_partials|k_ln_bwd16|k_global" -s $SKIP -c $CNT -f -o $OUT/${TAG}_prof_teacher_$cfg $CMD > $OUT/${TAG}_ncufull_run_$cfg.log 2>&1
  echo "ncu full $cfg exit=$?"
done
echo "gpu_round3 done"
