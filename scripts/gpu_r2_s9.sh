mkdir -p gpurun_out
python scripts/bseg_probe.py > gpurun_out/r2y_bseg_probe.txt 2>&1; cat gpurun_out/r2y_bseg_probe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_segments or every_call_path" > gpurun_out/r2y_pytest.log 2>&1; tail -4 gpurun_out/r2y_pytest.log
