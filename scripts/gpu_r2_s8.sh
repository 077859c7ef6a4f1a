mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_call_path" > gpurun_out/r2w_pytest.log 2>&1; tail -30 gpurun_out/r2w_pytest.log
