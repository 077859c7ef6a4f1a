mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python scripts/time_configs.py c1 c3 c4 c5_shard > gpurun_out/time_configs.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1
