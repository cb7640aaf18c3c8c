set -u
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/r2_pytest_a.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/r2_pytest_a.log
grep -E "parity\]|passed|failed|FAILED|Error" gpurun_out/r2_pytest_a.log | head -80
