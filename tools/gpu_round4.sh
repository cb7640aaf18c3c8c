#!/bin/bash
# Round-2 closing run on the final tree (one gpurun call, one GPU): full GPU test suite, smoke, the default bench line, the warm
# per-kernel list and the ncu launch list of the loss step, the host-buffer step's timeline.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r2h}
timeout 1500 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 1200 python bench.py > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench_1gpu.err; echo "bench exit=$?"
timeout 200 python tools/prof_loss_step.py 2>&1 | grep -v -i warn > $OUT/${TAG}_loss_step_warm_kernels.txt
timeout 200 python tools/trace_host_step.py 2>&1 | grep -v -i warn > $OUT/${TAG}_host_step_timeline.txt
CMD="python bench.py --steps 2 --warmup 3 --path tensor --no-retrieval --no-teacher --no-cpu-baseline --no-graph"
timeout 300 $CMD > $OUT/${TAG}_ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_loss_step.csv $CMD > $OUT/${TAG}_ncu_run.log 2>&1
echo "ncu launch list exit=$?"
echo "gpu_round4 done"
