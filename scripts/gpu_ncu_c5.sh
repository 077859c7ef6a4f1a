# ncu --set full of the three kernels of the BASELINE c5 shard (segmented program)
mkdir -p gpurun_out
FFTCONV_SKIP_REF=1 timeout 300 python scripts/time_configs.py c5_shard > gpurun_out/c5_plain.log 2>&1 && \
FFTCONV_SKIP_REF=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fc_fused_axis|fc_fast' -s 6 -c 3 -o gpurun_out/prof_c5 python scripts/time_configs.py c5_shard > gpurun_out/ncu_c5.log 2>&1
