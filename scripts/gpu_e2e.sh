mkdir -p gpurun_out
timeout 300 python scripts/e2e_timeline.py > gpurun_out/e2e_timeline.log 2>&1
