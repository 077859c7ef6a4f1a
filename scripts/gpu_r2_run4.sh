# Round 2, call 4: CTA de-synchronisation delay and kernel-spectrum preload before the phase barrier.
mkdir -p gpurun_out
O=gpurun_out/r2d_desync.txt
: > $O
python scripts/kb_probe.py c2 >> $O 2>&1
for d in 1000 2000 3000 4000 6000 8000 12000; do FFTCONV_B200_DESYNC=$d python scripts/kb_probe.py c2 >> $O 2>&1; done
for d in 2000 4000 6000; do FFTCONV_B200_DESYNC=$d FFTCONV_B200_PAIRKB=1,8,3 python scripts/kb_probe.py c2 >> $O 2>&1; done
python scripts/kb_probe.py c5 >> $O 2>&1
for d in 20000 50000 100000; do FFTCONV_B200_DESYNC=$d python scripts/kb_probe.py c5 >> $O 2>&1; done
python scripts/kb_probe.py img256 >> $O 2>&1
for d in 2000 4000; do FFTCONV_B200_DESYNC=$d python scripts/kb_probe.py img256 >> $O 2>&1; done
