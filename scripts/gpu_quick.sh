set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=600 -k "golden or baseline or modules" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'fc_fused_axis|fc_fast' -s 9 -c 3 -o gpurun_out/prof_c2_v3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
