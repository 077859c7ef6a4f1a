mkdir -p gpurun_out
python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'fc_fast|fc_plane|fc_contract' -s 10 -c 5 -o gpurun_out/prof_c3 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_c3.log 2>&1
