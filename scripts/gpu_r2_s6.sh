mkdir -p gpurun_out
O=gpurun_out/r2u_tcfused.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -x -q -k "tensor_core or transforms_next or batch_segments or c4 or tc_" > gpurun_out/r2u_pytest.log 2>&1; tail -5 gpurun_out/r2u_pytest.log >> $O
timeout 600 python scripts/time_configs.py c4 > gpurun_out/r2u_time_c4.log 2>&1; tail -3 gpurun_out/r2u_time_c4.log >> $O
