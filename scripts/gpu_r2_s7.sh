mkdir -p gpurun_out
O=gpurun_out/r2v_gemm_wide2.txt
: > $O
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q -k "tc_ or tensor_core or transforms_next or c4" > gpurun_out/r2v_pytest.log 2>&1; tail -3 gpurun_out/r2v_pytest.log >> $O
timeout 600 python scripts/time_configs.py c4 > gpurun_out/r2v_time_c4.log 2>&1; tail -2 gpurun_out/r2v_time_c4.log | cut -c1-1200 >> $O
timeout 300 python scripts/split_probe.py >> $O 2>&1
