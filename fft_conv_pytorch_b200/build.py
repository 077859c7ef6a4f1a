"""In-tree build of libfftconv_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m fft_conv_pytorch_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfftconv_b200.so")
SOURCES = ["fc_api.cu", "fc_plan.cpp"]
DEPS = ["fc_api.cu", "fc_plan.cpp", "fc_plan.h", "fc_types.h", "fc_kernels.cuh", "fc_fused.cuh", "fc_pair.cuh", "fc_column.cuh", "fc_plane.cuh", "fc_line.cuh", "fc_tune.h", "fc_tc.cuh", "fc_stream.cuh", os.path.join("..", "..", "include", "fftconv_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [
        _nvcc(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared",
        "-o", OUT,
    ]
    # The image's default g++ wrapper links libstdc++ statically, which breaks iostreams inside a dlopen()ed
    # library next to the system libstdc++ that numpy/torch load; the distro g++ links it dynamically.
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    if os.environ.get("FFTCONV_B200_PACKED", "1") != "0":
        cmd += ["-DFC_PACKED_F32X2"]  # FADD2 for complex add/sub (sm_100a packed fp32)
    if os.environ.get("FFTCONV_B200_TUNING", "0") == "1":
        cmd += ["-DFC_TUNING"]  # development build: timing-experiment knobs read from the environment (csrc/fc_tune.h)
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
