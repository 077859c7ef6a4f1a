# c2: K1/K4 variants for M = 256 (FFTCONV_B200_FAST = lines per warp, warps, CTAs per SM)
mkdir -p gpurun_out
L=gpurun_out/fast_sweep.log; : > $L
for v in "X=0" "FFTCONV_B200_FAST=2,8,4" "FFTCONV_B200_FAST=1,16,2" "X=1"; do
  echo "== c2 $v" >> $L
  env $v FFTCONV_SKIP_REF=1 python scripts/time_configs.py c2 >> $L 2>&1
done
FFTCONV_SKIP_REF=1 python scripts/time_configs.py c5_shard >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
