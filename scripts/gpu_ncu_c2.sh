# ncu --set full of the three kernels of BASELINE c2
mkdir -p gpurun_out
FFTCONV_SKIP_REF=1 timeout 300 python scripts/time_configs.py c2 > gpurun_out/c2_plain.log 2>&1 && \
FFTCONV_SKIP_REF=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fc_fused_axis|fc_fast' -s 6 -c 3 -o gpurun_out/prof_c2 python scripts/time_configs.py c2 > gpurun_out/ncu_c2.log 2>&1
