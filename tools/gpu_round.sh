#!/bin/bash
# One gpurun call: probes, parity tests, bench, launch list.  Everything is wrapped in `timeout`;
# logs go to gpurun_out/.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [stage ...]'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
STAGES="${*:-probe simt tensor bench ncu}"   # stages: probe simt tensor bench ncu ncufull
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1

has() { [[ " $STAGES " == *" $1 "* ]]; }

if has probe; then
  : > $OUT/probe.log
  timeout 60 build/bin/tc_probe tma >> $OUT/probe.log 2>&1; echo "exit=$?" >> $OUT/probe.log
  for v in 0 8 4 16 32 1 2 3 35 19; do
    timeout 60 build/bin/tc_probe mma $v >> $OUT/probe.log 2>&1; echo "exit=$?" >> $OUT/probe.log
  done
  cat $OUT/probe.log
fi

if has simt; then
  timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_loss_tensor_gpu.py \
      --ignore=tests/test_loss_tensor_gpu.py > $OUT/pytest_simt.log 2>&1
  echo "pytest simt exit=$?" | tee -a $OUT/pytest_simt.log
  tail -15 $OUT/pytest_simt.log
fi

TENSOR_OK=1
if has tensor; then
  timeout 600 python -m pytest tests/test_loss_tensor_gpu.py -q -m gpu > $OUT/pytest_tensor.log 2>&1
  TENSOR_OK=$?
  echo "pytest tensor exit=$TENSOR_OK" | tee -a $OUT/pytest_tensor.log
  tail -40 $OUT/pytest_tensor.log
fi

if has bench; then
  timeout 900 python bench.py --steps 10 --warmup 3 --path fp32 > $OUT/bench_fp32.json 2> $OUT/bench_fp32.err
  echo "bench fp32 exit=$?"; tail -3 $OUT/bench_fp32.err; cat $OUT/bench_fp32.json
  if [ "$TENSOR_OK" = "0" ]; then
    timeout 900 python bench.py --steps 20 --warmup 5 --path tensor --no-retrieval --no-cpu-baseline > $OUT/bench_tensor.json 2> $OUT/bench_tensor.err
    echo "bench tensor exit=$?"; tail -3 $OUT/bench_tensor.err; cat $OUT/bench_tensor.json
  fi
fi

if has ncu; then
  P=${NCU_PATH:-tensor}
  CMD="python bench.py --steps 2 --warmup 3 --path $P --no-retrieval --no-cpu-baseline --no-graph"
  timeout 600 $CMD > $OUT/ncu_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$P.csv $CMD > $OUT/ncu_run.log 2>&1
  echo "ncu exit=$?"; tail -3 $OUT/ncu_run.log
fi
if has ncufull; then
  # one full capture of the tensor-path kernels of the loss step (after the plain run passed) ...
  CMD="python bench.py --steps 1 --warmup 3 --path tensor --no-retrieval --no-cpu-baseline --no-graph"
  timeout 600 $CMD > $OUT/ncufull_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_infonce_bwd_tc|k_stream_tc" -s 6 -c 2 -f -o $OUT/prof_loss_tc $CMD > $OUT/ncufull_run.log 2>&1
  echo "ncufull loss exit=$?"; tail -2 $OUT/ncufull_run.log
  # ... and of the retrieval streaming kernel
  CMD2="python tools/prof_retrieval.py 1000000 16384"
  timeout 600 $CMD2 > $OUT/ncufull2_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_tc -s 2 -c 1 -f -o $OUT/prof_retrieval_tc $CMD2 > $OUT/ncufull2_run.log 2>&1
  echo "ncufull retrieval exit=$?"; tail -2 $OUT/ncufull2_run.log
fi
echo "gpu_round done"
