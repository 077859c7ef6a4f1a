mkdir -p gpurun_out
O=gpurun_out/r2p_stream.txt
: > $O
timeout 300 python scripts/stream_check.py >> $O 2>&1; echo "stream_check rc=$?" >> $O
for i in 1 2; do
timeout 120 python scripts/kb_probe.py c2 >> $O 2>&1
FFTCONV_B200_PROBE_FLAGS=4096 timeout 120 python scripts/kb_probe.py c2 >> $O 2>&1
done
timeout 300 python bench.py --quick --no-cpu-baseline --steps 200 > gpurun_out/r2p_bench.log 2>&1; echo "bench rc=$?" >> $O
