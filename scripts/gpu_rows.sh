# Conflict-free group-to-line assignment in K1 / K4 (short rows): GPU tests, per-config timing against the previous library.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_rows.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rows.log
cp fft_conv_pytorch_b200/libfftconv_b200.so variants/lib_rows.so
VARIANTS="prev rows prev rows" CONFIGS="c3 img128 img256 c5_shard" bash scripts/gpu_variants.sh
