mkdir -p gpurun_out
O=gpurun_out/r2s_bseg.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r2s_pytest_tc.log 2>&1; tail -3 gpurun_out/r2s_pytest_tc.log >> $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_segments or c4" > gpurun_out/r2s_pytest_bseg.log 2>&1; tail -5 gpurun_out/r2s_pytest_bseg.log >> $O
timeout 600 python scripts/time_configs.py c4 c1 > gpurun_out/r2s_time_c4.log 2>&1; tail -30 gpurun_out/r2s_time_c4.log >> $O
timeout 600 python scripts/time_configs.py --flags=128 c4 > gpurun_out/r2s_time_c4_noseg.log 2>&1; tail -12 gpurun_out/r2s_time_c4_noseg.log >> $O
