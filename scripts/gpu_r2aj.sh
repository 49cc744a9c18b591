#!/bin/bash
# Round-2 GPU visit AJ: the round's last commit -- tests, smoke, the driver's two bench command lines.
TAG=${1:-r2aj}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log; tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py --impl reference > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cut -c1-220 $O/${TAG}_bench_ref.json
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench.json')); print('cfg2', d['value']/1e6, d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, 'lockstep', d['e2e']['lockstep']['value']/1e6, 'roof', d['roofline']['frac'], 'reroot', d['roofline']['reroot']['frac'], 'cpu', d['cpu_baseline']['value']/1e6, 'ordinary', d['ordinary_launches']['value']/1e6, 'launches', d['gpu_launches'], d['clocks'])"; tail -3 $O/${TAG}_bench.err
