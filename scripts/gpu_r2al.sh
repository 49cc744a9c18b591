#!/bin/bash
# Round-2 GPU visit AL: bench with the per-workload launch mode table.
TAG=${1:-r2al}
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --named cfg3,cfg4,cfg5 --skip-cpu --steps 12 > $O/${TAG}_bench_named.json 2> $O/${TAG}_bench_named.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_named.json'))
print('cfg2', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, 'bits', d['config']['programmatic_bits'], d['config']['leaf_stand_in_cooperates'])
for n in d['config']['named']: print(n['name'], n['value']/1e6, 'M sims/s', n['ms_per_step'], 'ms e2e', n['e2e']['value']/1e6, 'lock', n['e2e']['lockstep']['value']/1e6, 'roof', n['roofline']['frac'], 'reroot', n['roofline']['reroot']['frac'], 'bits', n['config']['programmatic_bits'], n['config']['leaf_stand_in_cooperates'])"; tail -3 $O/${TAG}_bench_named.err
