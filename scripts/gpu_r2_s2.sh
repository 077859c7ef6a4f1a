O=gpurun_out/r2q_stream_micro3.txt
: > $O
for a in 0 1 2; do ./scripts/micro/sb_ns2_abl$a 3 1 0 >> $O 2>&1; ./scripts/micro/sb_ns3_abl$a 2 1 0 >> $O 2>&1; done
./scripts/micro/sb_ns3_abl0 1 1 0 >> $O 2>&1
./scripts/micro/sb_ns3_abl0 2 0 0 >> $O 2>&1
