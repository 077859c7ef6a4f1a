# Session-3 re-entry check: GPU tests, e2e probe, per-config timing, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1
FFTCONV_SKIP_REF=1 timeout 900 python scripts/time_configs.py > gpurun_out/time_configs.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
