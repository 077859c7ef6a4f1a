#!/bin/bash
# Memory and race checking of the kernel source on the host-thread emulation (see scripts/emul_families.py).
#   bash scripts/emul_sanitize.sh address|thread  -> profiles/r2_emul_<tool>sanitizer.log
set -u
TOOL=${1:-address}
CSRC=fft_conv_pytorch_b200/csrc
SO=/tmp/libfftconv_emul_${TOOL}.so
/usr/bin/g++ -O1 -g -std=c++20 -fPIC -pthread -DFC_CPU_EMUL -Itests/cpu_emul -I$CSRC -fsanitize=$TOOL -fno-omit-frame-pointer -Wno-unknown-pragmas \
  -shared -o $SO -x c++ $CSRC/fc_api.cu $CSRC/fc_plan.cpp tests/cpu_emul/cuda_shim.cpp || exit 1
if [ "$TOOL" = address ]; then RT=$(gcc -print-file-name=libasan.so); export ASAN_OPTIONS=detect_leaks=0; else RT=$(gcc -print-file-name=libtsan.so); export TSAN_OPTIONS="report_signal_unsafe=0 history_size=4"; fi
LOG=profiles/r2_emul_${TOOL}sanitizer.log
{ echo "# g++ -fsanitize=$TOOL build of the kernel source (tests/cpu_emul), scripts/emul_families.py"; FFTCONV_EMUL_SO=$SO LD_PRELOAD=$RT python scripts/emul_families.py "${@:2}"; echo "exit code $?"; } > $LOG 2>&1
tail -5 $LOG
