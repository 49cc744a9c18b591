"""The synthetic game + "network" that stands in for the user's pgx environment and JAX network (neither is installable in
this image): a device-side hash game (csrc/tz_synth.cu -> lib/libtz_synth.so, include/tz_synth.h) and its PyTorch wrapper
(synthetic.py).  Used by bench.py, the GPU tests, smoke() and the scripts; NOT part of the product package -- turbozero_b200
and libtz_b200.so never depend on it."""
