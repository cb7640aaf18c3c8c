#!/bin/bash
# Multi-GPU evidence run: usage  gpurun --gpus N -- 'bash tools/gpu_multi.sh N [tag]'   (parity worker, per-rank timeline, bench line)
set -u
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-r2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_mg${N}_gpus.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511 tests/mp/sharded_worker.py > $OUT/${TAG}_mg${N}_worker.log 2>&1
echo "worker exit=$?" | tee -a $OUT/${TAG}_mg${N}_worker.log; grep -v Warn $OUT/${TAG}_mg${N}_worker.log | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29512 tools/prof_sharded.py > $OUT/${TAG}_mg${N}_timeline.txt 2> $OUT/${TAG}_mg${N}_timeline.err
echo "timeline exit=$?"; grep -v "Warn\|warn" $OUT/${TAG}_mg${N}_timeline.txt | tail -16
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_FLAGS:-} > $OUT/${TAG}_mg${N}_bench.json 2> $OUT/${TAG}_mg${N}_bench.err
echo "bench exit=$?"; python - <<P
import json
try:
    l = json.loads(open("$OUT/${TAG}_mg${N}_bench.json").read().strip().splitlines()[-1])
    print("N=%d ms_per_step %.4f value %.3e parity %s weak %s" % (l["n_gpus"], l["ms_per_step"], l["value"], l.get("parity_check"), l.get("weak_scaling", {}).get("ms_per_step")))
    r = l.get("retrieval", {}).get("sharded_top10")
    if r: print("retrieval", r["ms"], r["queries_per_s"], r.get("fp32_recheck", {}).get("status"))
except Exception as e:
    print("bench parse failed", e)
P
