set -u
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_retrieval_gpu.py tests/test_retrieval_tensor_gpu.py -q -m gpu > gpurun_out/r2_pytest_b.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/r2_pytest_b.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_pytest_b.log | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-teacher --no-cpu-baseline --gallery 1000000 --queries 10000 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench exit=$?"; tail -2 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_b.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks','gpu_launches')})
print('roofline', d['roofline']['frac'], d['roofline']['avg_launch_us'])
print('flickr', d['retrieval']['flickr_eval'])
PY
