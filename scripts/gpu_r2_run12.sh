mkdir -p gpurun_out
for v in "default:0:128" "nopair:256:128" "ys64:0:64" "noystage:1024:128" "default2:0:128" "nopair2:256:128"; do
  n=${v%%:*}; r=${v#*:}; f=${r%%:*}; s=${r#*:}
  FFTCONV_B200_YSS=$s python bench.py --quick --no-cpu-baseline --steps 200 --plan-flags $f > gpurun_out/r2l_bench_$n.log 2>&1
done
