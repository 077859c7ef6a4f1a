# Round 2, call 5: new tests + new bench.py (all configs, GPU reference, strong-scaling records at N=1).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 --durations=8 > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2e_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/r2e_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2e_bench.log
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2e_bench_ref.log 2>&1
