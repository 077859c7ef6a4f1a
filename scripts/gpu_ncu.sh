mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'fc_fused_axis' -s 3 -c 1 -o gpurun_out/prof_kb_nb2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
