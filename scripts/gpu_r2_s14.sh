mkdir -p gpurun_out
python scripts/wide_probe.py "[((4, 1, 128, 128, 128), (4, 1, 9, 9, 9)), ((2, 8, 100, 100, 100), (8, 8, 5, 5, 5)), ((1, 4, 128, 128, 64), (4, 4, 7, 7, 7)), ((2, 16, 64, 64, 64), (16, 16, 7, 7, 7))]" > gpurun_out/r2f_plane128.txt 2>&1; cat gpurun_out/r2f_plane128.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
