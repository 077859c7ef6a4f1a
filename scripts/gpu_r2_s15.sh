mkdir -p gpurun_out
python scripts/wide_probe.py "[((64, 16, 1024), (16, 16, 33)), ((256, 64, 512), (64, 64, 9)), ((128, 8, 1000), (8, 8, 65)), ((1024, 4, 512), (4, 4, 17)), ((32, 32, 1024), (32, 32, 129))]" > gpurun_out/r2g_line.txt 2>&1; cat gpurun_out/r2g_line.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
