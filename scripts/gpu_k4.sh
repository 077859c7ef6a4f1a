# K4 vectorised lattice / bias stores: GPU tests, then per-config timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_k4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k4.log
FFTCONV_SKIP_REF=1 timeout 600 python scripts/time_configs.py c5_shard c2 img256 > gpurun_out/time_configs_k4.log 2>&1
