mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; tail -3 gpurun_out/r2i_pytest.log
C="[((32, 64, 56, 56), (64, 64, 3, 3)), ((16, 64, 128, 128), (64, 64, 5, 5)), ((256, 64, 512), (64, 64, 9)), ((16, 64, 65536), (64, 64, 4097)), ((8, 32, 256, 256), (32, 32, 9, 9)), ((16, 96, 65536), (96, 96, 4097))]"
python scripts/wide_probe.py "$C" > gpurun_out/r2i_paths.txt 2>&1; cat gpurun_out/r2i_paths.txt | cut -c1-120
