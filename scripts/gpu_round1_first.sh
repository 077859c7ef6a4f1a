set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 300 python bench.py --steps 10 --warmup 3 --config c1 --no-cpu-baseline > gpurun_out/bench_c1.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --config c3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1
timeout 600 python baseline/ref_probe.py cfg1_1d cfg2_2d cfg3_3d > gpurun_out/ref_probe_stdout.log 2>&1
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench.log
