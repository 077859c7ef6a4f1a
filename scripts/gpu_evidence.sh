# Evidence run (1 GPU): smoke, tests, bench (both arms), per-config timing, ncu launch list, ncu full capture of the c2 kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 900 python scripts/time_configs.py > gpurun_out/time_configs.log 2>&1
FFTCONV_SKIP_REF=1 timeout 600 python scripts/time_configs.py c5_full > gpurun_out/time_c5_full.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'fc_fused_axis|fc_fast' -s 9 -c 3 -o gpurun_out/prof_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
