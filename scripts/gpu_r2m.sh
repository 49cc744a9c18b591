#!/bin/bash
# Round-2 GPU visit M: SimP layout restored, walk-time L2 prefetch for the next launch (A/B against the same build without it
# and against the round-1 library, same box, alternating).
TAG=${1:-r2m}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -5 $O/${TAG}_pytest_gpu.log
for rep in 1 2 3; do
for lib in libtz_b200.so libtz_b200_nopf.so; do
  TZ_B200_LIB=$lib timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; ordinary', d['ordinary_launches']['value']/1e6)" | tee -a $O/${TAG}_variants.log
done
if [ -d scratch_r1 ]; then (cd scratch_r1 && timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('round-1 library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms')") | tee -a $O/${TAG}_variants.log; fi
done
for wl in cfg5 cfg1; do
for lib in libtz_b200.so libtz_b200_nopf.so; do
  TZ_B200_LIB=$lib timeout 300 python bench.py --workload $wl --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl $lib:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms')" | tee -a $O/${TAG}_variants.log
done; done
