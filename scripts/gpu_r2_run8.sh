# Round 2, call 8: the y-stage program (K1p/K4p with the radix-8 stage of the other axis, 64-point fused kernel).
mkdir -p gpurun_out
O=gpurun_out/r2h_ystage.txt
: > $O
python scripts/kb_probe.py c2 >> $O 2>&1
for oc in 4 3 2; do FFTCONV_B200_KB64OCC=$oc python scripts/kb_probe.py c2 >> $O 2>&1; done
FFTCONV_B200_FLAGSX=1 python scripts/kb_probe.py img256 >> $O 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log >> $O
python bench.py --quick --no-cpu-baseline --steps 100 > gpurun_out/r2h_bench.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fc_pair' -s 9 -c 3 -o gpurun_out/r2h_prof_ystage_c2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --quick > gpurun_out/ncu_full.log 2>&1
