# A/B of programmatic dependent launch (FFTCONV_B200_PDL=0 / 1): GPU tests with it on, bench and per-config timing both ways.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_pdl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pdl.log
for v in 0 1 0 1; do
  export FFTCONV_B200_PDL=$v
  timeout 300 python bench.py --no-cpu-baseline --steps 200 >> gpurun_out/bench_pdl$v.log 2>&1
done
for v in 0 1; do
  export FFTCONV_B200_PDL=$v
  FFTCONV_SKIP_REF=1 timeout 600 python scripts/time_configs.py > gpurun_out/time_configs_pdl$v.log 2>&1
done
