# Round 2, call 2: kernel-spectrum load variants of the pair fused kernel (16-byte loads, L1 prefetch, 2 items per thread).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log
export FFTCONV_SKIP_REF=1
run() { tag=$1; shift; env "$@" timeout 300 python scripts/time_configs.py c2 img256 c5_shard > gpurun_out/r2b_time_$tag.log 2>&1; }
run v1 X=1
run kpf1 FFTCONV_B200_KPF=1
run kpf2 FFTCONV_B200_KPF=2
run kpf4 FFTCONV_B200_KPF=4
run np2 FFTCONV_B200_PAIRKB=2,16,1
run np2kpf2 FFTCONV_B200_PAIRKB=2,16,1 FFTCONV_B200_KPF=2
run v1b X=1
