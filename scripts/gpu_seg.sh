# GPU parity + per-config timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
FFTCONV_SKIP_REF=1 timeout 900 python scripts/time_configs.py > gpurun_out/time_configs.log 2>&1
