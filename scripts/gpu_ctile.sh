mkdir -p gpurun_out
for t in "8,8,128,1" "8,8,32,4" "8,8,32,8" "16,4,32,4" "16,2,32,8" "16,2,64,4" "4,4,32,8"; do
  FFTCONV_B200_CTILE="$t" timeout 300 python scripts/time_configs.py c4 2>&1 | grep -o '"kernel": "contract", "ms": [0-9.]*' | sed "s/^/$t  /" >> gpurun_out/ctile.log
done
