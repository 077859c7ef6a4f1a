mkdir -p gpurun_out
for t in "" "4,4,128,1" "4,4,64,2" "4,8,64,1" "4,8,32,1" "4,8,256,1" "4,4,32,2" "8,8,128,1" "8,8,64,1" "16,4,128,1" "16,2,128,1" "16,4,64,1"; do
  if [ -n "$t" ]; then export FFTCONV_B200_CTILE="$t"; else unset FFTCONV_B200_CTILE; fi
  echo "VARIANT [$t]" >> gpurun_out/ctile.log
  timeout 300 python scripts/time_configs.py c3 img128 >> gpurun_out/ctile.log 2>&1
done
