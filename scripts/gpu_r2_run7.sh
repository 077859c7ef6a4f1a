# Round 2, call 7: does halving the kernel-spectrum traffic from L2 pay? (two pair items per CTA, shared through registers or L1)
mkdir -p gpurun_out
O=gpurun_out/r2g_kshare.txt
: > $O
python scripts/kb_probe.py c2 >> $O 2>&1
FFTCONV_B200_ABL=2 python scripts/kb_probe.py c2 >> $O 2>&1
for ks in 0 1; do for a in 0 2 32; do FFTCONV_B200_PAIRKB=2,16,1 FFTCONV_B200_KSHARE=$ks FFTCONV_B200_ABL=$a python scripts/kb_probe.py c2 >> $O 2>&1; done; done
./scripts/micro/l2bw > gpurun_out/r2_l2bw.txt 2>&1
