# c5 sharded over N GPUs + the N-GPU bench line
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/c5_sharded.py --steps 10 > gpurun_out/c5_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c5_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_${N}gpu.log
