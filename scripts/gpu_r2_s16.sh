mkdir -p gpurun_out
C="[((32, 64, 56, 56), (64, 64, 3, 3)), ((16, 64, 128, 128), (64, 64, 5, 5)), ((256, 64, 512), (64, 64, 9)), ((16, 64, 65536), (64, 64, 4097)), ((8, 32, 256, 256), (32, 32, 9, 9)), ((32, 32, 1024), (32, 32, 129)), ((16, 96, 65536), (96, 96, 4097)), ((8, 128, 64, 64), (128, 128, 7, 7)), ((32, 24, 128, 128), (48, 24, 3, 3))]"
echo "--- default" > gpurun_out/r2h_ctiled.txt
python scripts/wide_probe.py "$C" >> gpurun_out/r2h_ctiled.txt 2>&1
echo "--- FC_FLAG_NO_TC" >> gpurun_out/r2h_ctiled.txt
python scripts/wide_probe.py "$C" 32 >> gpurun_out/r2h_ctiled.txt 2>&1
cat gpurun_out/r2h_ctiled.txt
