mkdir -p gpurun_out
FFTCONV_SKIP_REF=1 python scripts/time_configs.py c3 > gpurun_out/plain_c3.log 2>&1 && \
FFTCONV_SKIP_REF=1 ncu --set full --clock-control none --import-source on -k regex:'fc_pass_kernel' -s 30 -c 6 -o gpurun_out/prof_c3 python scripts/time_configs.py c3 > gpurun_out/ncu_c3.log 2>&1
