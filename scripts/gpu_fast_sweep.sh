mkdir -p gpurun_out
L=gpurun_out/fast_sweep.log; : > $L
for v in "X=0" "FFTCONV_B200_KB_L1PF=1" "X=1" "FFTCONV_B200_KB_L1PF=1"; do
  echo "== c2 $v" >> $L
  env $v FFTCONV_SKIP_REF=1 python scripts/time_configs.py c2 >> $L 2>&1
done
