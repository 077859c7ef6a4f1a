mkdir -p gpurun_out
for d in 0 1 2; do FFTCONV_B200_DBG=$d timeout 300 python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench_dbg$d.log 2>&1; done
