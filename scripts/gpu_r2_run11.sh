mkdir -p gpurun_out
O=gpurun_out/r2k_k1_ablation.txt
: > $O
for a in 0 1 2 4 3 5 6 7; do FFTCONV_B200_ABL1=$a python scripts/kb_probe.py c2 >> $O 2>&1; done
echo "--- no y stage (flags 1024)" >> $O
for a in 0 1 2 4 7; do FFTCONV_B200_ABL1=$a FFTCONV_B200_PROBE_FLAGS=1024 python scripts/kb_probe.py c2 >> $O 2>&1; done
echo "--- y stage, YSS=64" >> $O
for a in 0 4 7; do FFTCONV_B200_ABL1=$a FFTCONV_B200_YSS=64 python scripts/kb_probe.py c2 >> $O 2>&1; done
