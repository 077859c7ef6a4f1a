mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2b_bench_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_bench_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_bench_ref_2gpu.log
tail -c 600 gpurun_out/r2b_bench_2gpu.log
