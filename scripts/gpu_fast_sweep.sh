mkdir -p gpurun_out
L=gpurun_out/fast_sweep.log; : > $L
echo "== write" >> $L
python bench.py --no-cpu-baseline --flush write >> $L 2>&1
echo "== write+read" >> $L
python bench.py --no-cpu-baseline --flush write+read >> $L 2>&1
echo "== write+read c1" >> $L
python bench.py --no-cpu-baseline --config c1 >> $L 2>&1
echo "== write+read c3" >> $L
python bench.py --no-cpu-baseline --config c3 >> $L 2>&1
