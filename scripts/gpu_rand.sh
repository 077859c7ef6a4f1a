mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=900 -k "random_" --durations=5 > gpurun_out/pytest_rand.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_rand.log
