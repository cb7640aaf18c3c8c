set -u
cd /root/repo
mkdir -p gpurun_out
DCLIP_B200_LIB=build/prof/libdclip_b200_prof.so timeout 120 python tools/prof_gemm.py 2>&1 | tail -18
timeout 300 python tools/bench_gemm.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_xattn_gpu.py tests/test_optim_assembly_gpu.py -q -m gpu -s > gpurun_out/r2_pytest_b.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/r2_pytest_b.log
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r2_pytest_b.log | head -40
timeout 300 python tools/bench_xattn.py bf16 2>&1 | tail -1
