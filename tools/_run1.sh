set -u
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_xattn_gpu.py tests/test_loss_gpu.py tests/test_optim_assembly_gpu.py -q -m gpu > gpurun_out/r2_pytest_b.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/r2_pytest_b.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_pytest_b.log | head -20
timeout 300 python tools/prof_teacher_step.py bf16 > gpurun_out/r2_teacher_warm_bf16_b.txt 2>&1; echo "prof exit=$?"
cat gpurun_out/r2_teacher_warm_bf16_b.txt | grep -v Warn | head -20
timeout 300 python tools/bench_xattn.py bf16 2>&1 | tail -1
