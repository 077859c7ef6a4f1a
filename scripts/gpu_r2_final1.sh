# Round 2 evidence run (1 GPU, product build): tests, smoke, bench (both arms), ncu launch list, ncu full captures.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 --durations=5 > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_final_smoke.log
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2_final_bench_ref.log 2>&1
( time timeout 1500 python bench.py ) > gpurun_out/r2_final_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_final_bench.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fc_fused_axis|fc_fast' -s 9 -c 3 -o /tmp/prof_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_c2.ncu-rep --page raw --csv > gpurun_out/r2_final_ncu_c2_raw.csv 2>/dev/null
ncu -i /tmp/prof_c2.ncu-rep --page source --csv --kernel-name regex:fc_fused_axis > gpurun_out/r2_final_ncu_c2_fused_source.csv 2>/dev/null
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --quick --plan-flags 512 > gpurun_out/plain3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fc_pair' -s 9 -c 3 -o /tmp/prof_c2_ystage python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --quick --plan-flags 512 > gpurun_out/ncu_full2.log 2>&1
ncu -i /tmp/prof_c2_ystage.ncu-rep --page raw --csv > gpurun_out/r2_final_ncu_c2_ystage_raw.csv 2>/dev/null
timeout 600 python scripts/time_configs.py c1 c2 c3 img128 img256 c5_shard > gpurun_out/r2_final_time_configs.log 2>&1
