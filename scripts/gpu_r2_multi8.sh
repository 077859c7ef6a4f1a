mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > gpurun_out/r2b_bench_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_bench_8gpu.log
tail -c 900 gpurun_out/r2b_bench_8gpu.log
