mkdir -p gpurun_out; rm -f gpurun_out/tune.log
cp fft_conv_pytorch_b200/libfftconv_b200.so /tmp/default.so
for v in default pf8 pf16 pf32; do
  if [ $v = default ]; then cp /tmp/default.so fft_conv_pytorch_b200/libfftconv_b200.so; else cp fft_conv_pytorch_b200/libfftconv_b200_$v.so fft_conv_pytorch_b200/libfftconv_b200.so; fi
  echo "VARIANT $v" >> gpurun_out/tune.log
  FFTCONV_SKIP_REF=1 timeout 300 python scripts/time_configs.py c2 img256 c5_shard >> gpurun_out/tune.log 2>&1
done
