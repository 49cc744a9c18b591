#!/bin/bash
# Diagnostic build of the libraries with in-kernel phase clocks (-DTZ_PROFILE): turbozero_b200/lib/libtz_b200_prof.so and
# libtz_synth_prof.so.  Used by scripts/phase_*.py and scripts/timeline.py only; never loaded by the product path, the tests
# or bench.py.
set -e
cd "$(dirname "$0")/.."
python -m turbozero_b200.build --prof "$@"
python -m standin.build --prof "$@"
