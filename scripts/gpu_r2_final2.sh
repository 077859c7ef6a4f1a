# Round 2, second evidence run (1 GPU, product build): tests, smoke, bench (both arms), ncu launch list, ncu full captures.
mkdir -p gpurun_out
P=gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 --durations=5 > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?" >> ${P}_smoke.log
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > ${P}_bench_ref.log 2>&1
( time timeout 1500 python bench.py ) > ${P}_bench.log 2>&1; echo "bench rc=$?" >> ${P}_bench.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fc_fused_axis|fc_fast|fc_stream' -s 9 -c 3 -o /tmp/prof_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_c2.ncu-rep --page raw --csv > ${P}_ncu_c2_raw.csv 2>/dev/null
timeout 600 python scripts/time_configs.py c1 c2 c3 c4 img128 img256 c5_shard > ${P}_time_configs.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fc_tc_|fc_col_' -s 7 -c 5 -o /tmp/prof_c4 python scripts/time_configs.py c4 > gpurun_out/ncu_full_c4.log 2>&1
ncu -i /tmp/prof_c4.ncu-rep --page raw --csv > ${P}_ncu_c4_raw.csv 2>/dev/null
cuobjdump -sass fft_conv_pytorch_b200/libfftconv_b200.so 2>/dev/null | awk '/Function : /{f=$3} /UBLKCP|UTMALDG|UTMASTG|UTCHMMA|UTCBAR|LDTM|SYNCS/{c[f" "$2]++} END{for(k in c) print c[k], k}' | sort -k2 > ${P}_sass_async_mnemonics.txt
