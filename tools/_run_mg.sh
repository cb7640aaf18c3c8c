set -u
cd /root/repo
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_mg${N}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -s > gpurun_out/r2_mg${N}_pytest.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/r2_mg${N}_pytest.log
grep -E "sharded parity|SHARDED_OK|passed|failed|skipped|Error" gpurun_out/r2_mg${N}_pytest.log | head -60
for W in 2 4 8; do
  if [ $W -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2961$W bench.py --gpus $W --steps 20 --warmup 5 > gpurun_out/r2_mg${W}_bench.json 2> gpurun_out/r2_mg${W}_bench.err; echo "bench N=$W exit=$?"
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2962$W tools/prof_sharded.py > gpurun_out/r2_mg${W}_timeline.txt 2> gpurun_out/r2_mg${W}_timeline.err; echo "timeline N=$W exit=$?"
  fi
done
