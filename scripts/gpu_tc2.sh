mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q --timeout=200 -k "tc or c4" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/time_configs.py c4 > gpurun_out/time_configs.log 2>&1
