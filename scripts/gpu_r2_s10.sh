mkdir -p gpurun_out
O=gpurun_out/r2z_c1.txt
: > $O
python scripts/time_configs.py c1 2>&1 | tail -1 | cut -c1-700 >> $O
python scripts/time_configs.py --flags=16384 c1 2>&1 | tail -1 | cut -c1-700 >> $O
