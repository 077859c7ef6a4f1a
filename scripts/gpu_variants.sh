# A/B of prebuilt library variants (variants/lib_<name>.so): per-config timing and the bench, alternating.
mkdir -p gpurun_out
for v in $VARIANTS; do
  cp variants/lib_$v.so fft_conv_pytorch_b200/libfftconv_b200.so
  echo "VARIANT $v" >> gpurun_out/variants.log
  FFTCONV_SKIP_REF=1 timeout 600 python scripts/time_configs.py $CONFIGS >> gpurun_out/variants.log 2>&1
  timeout 300 python bench.py --no-cpu-baseline --steps 200 >> gpurun_out/variants_bench_$v.log 2>&1
done
