mkdir -p gpurun_out
for c in 1 4 5 6 8; do
FFTCONV_B200_HOST_CHUNKS=$c timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_short$c.log 2>&1
done
