# ncu --set full of the kernels added in the second half of round 2 that the c2 / c4 captures do not contain
mkdir -p gpurun_out
C="[((256, 64, 512), (64, 64, 9)), ((16, 64, 65536), (64, 64, 4097)), ((4, 1, 128, 128, 128), (4, 1, 9, 9, 9)), ((32, 32, 8192), (64, 32, 129))]"
python scripts/wide_probe.py "$C" > gpurun_out/plain_misc.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fc_line_|fc_contract_tiled|fc_plane_|fc_col_|fc_fast_c2c' -s 40 -c 14 -o /tmp/prof_misc python scripts/wide_probe.py "$C" > gpurun_out/ncu_misc.log 2>&1
ncu -i /tmp/prof_misc.ncu-rep --page raw --csv > gpurun_out/r2b_ncu_misc_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_misc.log
