# Round 2, first GPU call: the packed batch-pair pipeline on real hardware (tests, timing A/B, ncu, sanitizer).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 --durations=5 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
timeout 300 python scripts/sanitize_families.py > gpurun_out/r2_families.log 2>&1; echo "rc=$?" >> gpurun_out/r2_families.log
export FFTCONV_SKIP_REF=1
timeout 600 python scripts/time_configs.py c1 c2 c3 img128 img256 c5_shard > gpurun_out/r2_time_pair.log 2>&1
timeout 600 python scripts/time_configs.py --flags=256 c2 img256 c5_shard > gpurun_out/r2_time_nopair.log 2>&1
for v in "2,2"; do FFTCONV_B200_PAIRROW=$v timeout 300 python scripts/time_configs.py c2 c5_shard > gpurun_out/r2_time_pairrow_$v.log 2>&1; done
timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/r2_bench_c2.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fc_pair' -s 9 -c 3 -o gpurun_out/r2_prof_pair_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_families.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 3 python scripts/sanitize_families.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log
