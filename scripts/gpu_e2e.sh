mkdir -p gpurun_out
timeout 300 python scripts/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1
