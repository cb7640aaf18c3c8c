#!/bin/bash
# Round-2 evidence run (one gpurun call): full GPU test suite, the default bench line, the ncu launch list of the loss step and
# one --set full capture of its three tensor kernels.  Logs / reports land in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r2}
timeout 1500 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 1200 python bench.py > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err; echo "bench exit=$?"
tail -2 $OUT/${TAG}_bench_1gpu.err
CMD="python bench.py --steps 2 --warmup 3 --path tensor --no-retrieval --no-teacher --no-cpu-baseline --no-graph"
timeout 600 $CMD > $OUT/${TAG}_ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_loss_step.csv $CMD > $OUT/${TAG}_ncu_run.log 2>&1
echo "ncu launch list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_pair|k_stream_tc" -s 9 -c 3 -f -o $OUT/${TAG}_prof_loss_kernels $CMD > $OUT/${TAG}_ncufull_run.log 2>&1
echo "ncu full exit=$?"; tail -2 $OUT/${TAG}_ncufull_run.log
echo "gpu_round2 done"
