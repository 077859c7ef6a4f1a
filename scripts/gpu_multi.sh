# N-GPU bench under torchrun (both arms), N from $1
N=${1:-4}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_${N}gpu.log
