# Bank-disjoint line assignment (K1 / K4 short rows, plane kernels): GPU tests, per-config timing against the previous library.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_rows.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rows.log
cp fft_conv_pytorch_b200/libfftconv_b200.so variants/lib_new.so
VARIANTS="prev new prev new" CONFIGS="c3" bash scripts/gpu_variants.sh
cp variants/lib_new.so fft_conv_pytorch_b200/libfftconv_b200.so
