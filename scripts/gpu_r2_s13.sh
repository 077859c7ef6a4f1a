mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r2e_pytest_tc.log 2>&1; tail -4 gpurun_out/r2e_pytest_tc.log
python scripts/wide_probe.py "[((32, 64, 56, 56), (64, 64, 3, 3)), ((16, 64, 128, 128), (64, 64, 5, 5)), ((256, 64, 512), (64, 64, 9)), ((16, 64, 65536), (64, 64, 4097)), ((8, 32, 256, 256), (64, 32, 9, 9)), ((16, 96, 65536), (192, 96, 4097))]" > gpurun_out/r2e_wide_probe.txt 2>&1; cat gpurun_out/r2e_wide_probe.txt
