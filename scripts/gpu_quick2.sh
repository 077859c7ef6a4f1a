mkdir -p gpurun_out
FFTCONV_SKIP_REF=1 timeout 900 python scripts/time_configs.py c2 c5_shard img256 c1 > gpurun_out/time_configs.log 2>&1
