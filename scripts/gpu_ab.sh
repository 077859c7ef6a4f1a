mkdir -p gpurun_out
for t in "nb=1,occ=3" "nb=1,occ=4" "nb=2,occ=3" "nb=1,occ=2"; do
  export FFTCONV_B200_TUNE="$t"
  timeout 300 python -m pytest tests -m gpu -x -q --timeout=300 -k "baseline_c2 or modules" > gpurun_out/pytest_$t.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$t.log
  timeout 300 python bench.py --no-cpu-baseline --steps 100 > "gpurun_out/bench_$t.log" 2>&1
done
