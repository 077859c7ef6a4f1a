mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/time_configs.py c1 c4 > gpurun_out/time_configs.log 2>&1
timeout 200 python scripts/tc_bench.py 8192 > gpurun_out/tc_bench.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'fc_tc_gemm' -s 2 -c 1 -o gpurun_out/prof_tc python scripts/tc_bench.py 8192 > gpurun_out/ncu_tc.log 2>&1
