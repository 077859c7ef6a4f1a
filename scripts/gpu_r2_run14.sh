mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/r2n_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_gpu.log
timeout 300 python scripts/time_configs.py c4 > gpurun_out/r2n_time_c4.log 2>&1
