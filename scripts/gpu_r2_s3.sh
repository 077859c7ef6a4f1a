mkdir -p gpurun_out
O=gpurun_out/r2r_stream_lib.txt
: > $O
for i in 1 2; do
timeout 120 python scripts/kb_probe.py c2 >> $O 2>&1
FFTCONV_B200_PROBE_FLAGS=4096 timeout 120 python scripts/kb_probe.py c2 >> $O 2>&1
FFTCONV_B200_PROBE_FLAGS=8192 timeout 120 python scripts/kb_probe.py c2 >> $O 2>&1
done
timeout 300 python bench.py --quick --no-cpu-baseline --steps 200 > gpurun_out/r2r_bench.log 2>&1; echo "bench rc=$?" >> $O
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; tail -2 gpurun_out/r2r_pytest.log >> $O
