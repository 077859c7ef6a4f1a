# c4 / c3: tile budget of the transposing generic passes (FFTCONV_B200_TILE_RFAST) and threads.
mkdir -p gpurun_out
L=gpurun_out/fast_sweep.log; : > $L
for v in "X=0" "FFTCONV_B200_TILE_RFAST=4096" "FFTCONV_B200_TILE_RFAST=8192" "FFTCONV_B200_TILE_RFAST=4096 FFTCONV_B200_THREADS=512" "FFTCONV_B200_TILE_RFAST=1024"; do
  echo "== $v" >> $L
  env $v FFTCONV_SKIP_REF=1 python scripts/time_configs.py c4 c3 c1 >> $L 2>&1
done
