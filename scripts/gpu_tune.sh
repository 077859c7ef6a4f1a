mkdir -p gpurun_out; rm -f gpurun_out/tune.log
echo "VARIANT default" >> gpurun_out/tune.log
FFTCONV_SKIP_REF=1 timeout 300 python scripts/time_configs.py c2 img256 c1 >> gpurun_out/tune.log 2>&1
cp fft_conv_pytorch_b200/libfftconv_b200_vec2.so fft_conv_pytorch_b200/libfftconv_b200.so
echo "VARIANT vec2" >> gpurun_out/tune.log
FFTCONV_SKIP_REF=1 timeout 300 python scripts/time_configs.py c2 img256 c1 >> gpurun_out/tune.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2 or golden" >> gpurun_out/tune.log 2>&1
