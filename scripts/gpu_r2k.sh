#!/bin/bash
# Round-2 GPU visit K: host-built re-root descriptors; where the 2 % of the headline went (experiment builds without the
# optional timeline record / without the q_transform switch, and the round-1 library, all on the same box).
TAG=${1:-r2k}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -5 $O/${TAG}_pytest_gpu.log
for wl in cfg2 cfg4 cfg3 cfg5; do
  timeout 600 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 4 2>$O/${TAG}_$wl.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; launch', r['avg_launch_us'], 'us frac', r['frac'], '; reroot', r['reroot']['avg_launch_us'], 'us frac', r['reroot']['frac'])" | tee -a $O/${TAG}_bench.log
done
for rep in 1 2; do
for lib in libtz_b200.so libtz_b200_notl.so libtz_b200_noqt.so; do
  TZ_B200_LIB=$lib timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; ordinary', d['ordinary_launches']['value']/1e6)" | tee -a $O/${TAG}_variants.log
done
if [ -d scratch_r1 ]; then (cd scratch_r1 && timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('round-1 library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; ordinary', d['config']['ordinary_launches']['value']/1e6)") | tee -a $O/${TAG}_variants.log; fi
done
timeout 300 python scripts/phase_r2.py reroot connect_four 1024 128 256 1 > $O/${TAG}_phase_reroot_c4.log 2>&1; tail -14 $O/${TAG}_phase_reroot_c4.log
