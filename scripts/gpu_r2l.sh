#!/bin/bash
# Round-2 GPU visit L: compile-time q_transform / mode-bit timeline in k_sim (headline back to round 1's level?), full bench.
TAG=${1:-r2l}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -5 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
for rep in 1 2; do
timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('this library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms; ordinary', d['ordinary_launches']['value']/1e6)" | tee -a $O/${TAG}_variants.log
if [ -d scratch_r1 ]; then (cd scratch_r1 && timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('round-1 library:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'ms')") | tee -a $O/${TAG}_variants.log; fi
done
timeout 600 python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; cat $O/${TAG}_bench_cfg2.json | cut -c1-1500; tail -5 $O/${TAG}_bench_cfg2.err
timeout 900 python bench.py --named cfg3,cfg4,cfg5 --skip-cpu --steps 8 > $O/${TAG}_bench_named.json 2> $O/${TAG}_bench_named.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_named.json'))
for n in d['config']['named']: print(n['name'], n['value']/1e6, 'M sims/s', n['ms_per_step'], 'ms e2e', n['e2e']['value']/1e6, 'roof', n['roofline']['frac'], 'reroot', n['roofline']['reroot']['frac'])"; tail -5 $O/${TAG}_bench_named.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cut -c1-400 $O/${TAG}_bench_ref.json
