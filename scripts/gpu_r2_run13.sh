mkdir -p gpurun_out
O=gpurun_out/r2m_c5_fill.txt
: > $O
python scripts/kb_probe.py c5 7 >> $O 2>&1
FFTCONV_B200_PROBE_FLAGS=2048 python scripts/kb_probe.py c5 7 >> $O 2>&1
python scripts/kb_probe.py c5 7 >> $O 2>&1
python scripts/kb_probe.py c2 >> $O 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; tail -2 gpurun_out/r2m_pytest.log >> $O
