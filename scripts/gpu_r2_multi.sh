# torchrun bench on N GPUs of one box (strong-scaling records + weak headline), N from $1
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2_final_bench_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_bench_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_final_bench_ref_${N}gpu.log 2>&1
