# c5 shard: generic-pass knobs (flat tile budget, threads) after the K1/K4 rewrite.
mkdir -p gpurun_out
L=gpurun_out/fast_sweep.log; : > $L
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for v in "X=0" "FFTCONV_B200_TILE_FLAT=2048" "FFTCONV_B200_TILE_FLAT=8192" "FFTCONV_B200_TILE_FLAT=16384" "FFTCONV_B200_THREADS=512" "FFTCONV_B200_THREADS=128" "FFTCONV_B200_THREADS=512 FFTCONV_B200_TILE_FLAT=8192" "FFTCONV_B200_THREADS=512 FFTCONV_B200_TILE_FLAT=2048"; do
  echo "== c5 $v" >> $L
  env $v FFTCONV_SKIP_REF=1 python scripts/time_configs.py c5_shard >> $L 2>&1
done
