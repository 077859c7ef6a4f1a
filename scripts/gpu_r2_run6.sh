# Round 2, call 6: two warps per pair line (16-warp CTAs at 64 registers) against one warp per line.
mkdir -p gpurun_out
O=gpurun_out/r2f_tpl64.txt
: > $O
for i in 1 2; do
python scripts/kb_probe.py c2 >> $O 2>&1
FFTCONV_B200_PAIRKB=1,16,2 python scripts/kb_probe.py c2 >> $O 2>&1
done
FFTCONV_B200_PAIRKB=1,16,2 FFTCONV_B200_KPF=2 python scripts/kb_probe.py c2 >> $O 2>&1
for a in 2 32 1 4 16 63; do FFTCONV_B200_PAIRKB=1,16,2 FFTCONV_B200_ABL=$a python scripts/kb_probe.py c2 >> $O 2>&1; done
python bench.py --quick --no-cpu-baseline --steps 50 > gpurun_out/r2f_bench_default.log 2>&1
FFTCONV_B200_PAIRKB=1,16,2 python bench.py --quick --no-cpu-baseline --steps 50 > gpurun_out/r2f_bench_w16.log 2>&1
FFTCONV_B200_PAIRKB=1,16,2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair or c2 or golden" > gpurun_out/r2f_pytest_w16.log 2>&1
