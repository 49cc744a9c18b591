#!/bin/bash
# Diagnostic build of the search library with in-kernel phase clocks (-DTZ_PROFILE): turbozero_b200/lib/libtz_b200_prof.so
# Used by scripts/phase_clocks.py only; never loaded by the product path, the tests or bench.py.
set -e
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -DTZ_PROFILE \
  -Iinclude turbozero_b200/csrc/tz_kernels.cu turbozero_b200/csrc/tz_replay.cu -o turbozero_b200/lib/libtz_b200_prof.so
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -DTZ_PROFILE \
  -Iinclude turbozero_b200/csrc/tz_synth.cu -o turbozero_b200/lib/libtz_synth_prof.so
