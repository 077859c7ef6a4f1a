# c2: fused axis kernel with 4 batches per CTA (16 warps) vs the default; c3/c5 with the 4x8 contraction rule.
mkdir -p gpurun_out
L=gpurun_out/fast_sweep.log; : > $L
for v in "X=0" "FFTCONV_B200_TUNE=nb=4" "X=1" "FFTCONV_B200_TUNE=nb=4"; do
  echo "== c2 $v" >> $L
  env $v FFTCONV_SKIP_REF=1 python scripts/time_configs.py c2 >> $L 2>&1
done
echo "== c3 c5" >> $L
FFTCONV_SKIP_REF=1 python scripts/time_configs.py c3 c5_shard >> $L 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
