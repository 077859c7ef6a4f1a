# Round 2, call 3: phase ablation of the pair fused kernel (tuning build) + occupancy variants.
mkdir -p gpurun_out
O=gpurun_out/r2c_ablation.txt
: > $O
for a in 0 1 2 4 8 16 32 3 5 6 7 10 15 31 63 47 34 33 36; do FFTCONV_B200_ABL=$a python scripts/kb_probe.py c2 >> $O 2>&1; done
for v in "1,8,3" "1,4,3" "2,16,1"; do FFTCONV_B200_PAIRKB=$v python scripts/kb_probe.py c2 >> $O 2>&1; done
for a in 0 2 32 63; do FFTCONV_B200_ABL=$a python scripts/kb_probe.py c5 >> $O 2>&1; done
FFTCONV_B200_PDL=0 python scripts/kb_probe.py c2 >> $O 2>&1
