mkdir -p gpurun_out
O=gpurun_out/r2i_ystage.txt
: > $O
python scripts/kb_probe.py c2 >> $O 2>&1
python scripts/kb_probe.py c2 >> $O 2>&1
python scripts/kb_probe.py img256 >> $O 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; tail -2 gpurun_out/r2i_pytest.log >> $O
python bench.py --quick --no-cpu-baseline --steps 100 > gpurun_out/r2i_bench.log 2>&1
