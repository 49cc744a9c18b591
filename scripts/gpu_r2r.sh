#!/bin/bash
# Round-2 GPU visit R: after the stand-in moved to standin/ and the e2e leg became an actor loop -- full check.
TAG=${1:-r2r}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_cfg2.json')); print('cfg2', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, 'lockstep', d['e2e']['lockstep']['value']/1e6, 'roof', d['roofline']['frac'], 'reroot', d['roofline']['reroot']['frac'], 'cpu', d['cpu_baseline']['value']/1e6, d['cpu_baseline']['numpy_port']['value'])"; tail -3 $O/${TAG}_bench_cfg2.err
timeout 900 python bench.py --named cfg3,cfg4,cfg5 --skip-cpu --steps 8 > $O/${TAG}_bench_named.json 2> $O/${TAG}_bench_named.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_named.json'))
for n in d['config']['named']: print(n['name'], n['value']/1e6, 'M sims/s', n['ms_per_step'], 'ms e2e', n['e2e']['value']/1e6, 'lock', n['e2e']['lockstep']['value']/1e6, 'roof', n['roofline']['frac'], 'reroot', n['roofline']['reroot']['frac'])"; tail -3 $O/${TAG}_bench_named.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cut -c1-200 $O/${TAG}_bench_ref.json
timeout 200 python scripts/phase_r2.py wide othello 512 200 400 2 weighted 2>&1 | tail -12
