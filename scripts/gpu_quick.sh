mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=600 -k "golden or baseline or modules" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
