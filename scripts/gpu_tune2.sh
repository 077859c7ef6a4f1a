mkdir -p gpurun_out
for t in "" "pgrid=2" "pgrid=4" "" "pgrid=2" "pgrid=3"; do
  export FFTCONV_B200_TUNE="$t"
  echo "VARIANT [$t]" >> gpurun_out/tune2.log
  FFTCONV_SKIP_REF=1 timeout 600 python scripts/time_configs.py c2 img256 c5_shard c1 >> gpurun_out/tune2.log 2>&1
done
