mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
FFTCONV_B200_HOST_CHUNKS=1 timeout 600 python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench_chunks1.log 2>&1
FFTCONV_B200_HOST_CHUNKS=8 timeout 600 python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench_chunks8.log 2>&1
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu.log
