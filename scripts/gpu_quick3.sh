mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_ct.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ct.log
timeout 300 python scripts/time_configs.py c3 img128 c1 > gpurun_out/time_ct.log 2>&1
