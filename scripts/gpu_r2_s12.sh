mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -4 gpurun_out/r2d_pytest.log
python scripts/short_split_probe.py > gpurun_out/r2d_short_split_probe.txt 2>&1; cat gpurun_out/r2d_short_split_probe.txt
python scripts/bseg_probe.py > gpurun_out/r2d_bseg_probe.txt 2>&1; cat gpurun_out/r2d_bseg_probe.txt
