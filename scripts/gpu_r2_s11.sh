mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; tail -4 gpurun_out/r2c_pytest.log
python scripts/tiny_probe.py > gpurun_out/r2c_tiny_probe.txt 2>&1
python scripts/time_configs.py c1 2>&1 | tail -1 | cut -c1-400
timeout 900 python scripts/kernel_size_sweep.py > gpurun_out/r2c_sweep.log 2>&1; tail -3 gpurun_out/r2c_sweep.log | cut -c1-300
