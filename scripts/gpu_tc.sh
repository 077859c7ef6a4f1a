mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_tc.py -x -q -s --timeout=120 > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tc.log
