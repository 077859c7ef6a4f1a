mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q -k "tc or c4 or matmul" > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tc.log
FFTCONV_SKIP_REF=1 timeout 600 python scripts/time_configs.py c4 > gpurun_out/time_c4.log 2>&1
timeout 300 python scripts/tc_bench.py 8192 > gpurun_out/tc_bench.log 2>&1
