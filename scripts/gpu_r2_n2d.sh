#!/bin/bash
# Round-2 two-GPU visit D: the driver's command line at N = 2 with the final bench (launch-mode table), plus configs[4] forced
# (its NCCL legs) and the cross-rank replay sample check.
TAG=${1:-r2n2d}
O=gpurun_out
mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu2_replay_sample_check.py > $O/${TAG}_sample_check.log 2>&1; echo "rc=$?" >> $O/${TAG}_sample_check.log; tail -3 $O/${TAG}_sample_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 12 --warmup 4 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -2 $O/${TAG}_bench.err | cut -c1-300; python -c "
import json; d=json.load(open('$O/${TAG}_bench.json')); print('N=2:', d['value']/1e6, 'M sims/s e2e', d['e2e']['value']/1e6, d['clocks'])
for n in d['config']['named']: print(' named', n['name'], n['scaling'], n['config']['envs_per_gpu'], 'envs/gpu', n['value']/1e6, 'M sims/s', n['ms_per_step'], 'ms e2e', n['e2e']['value']/1e6)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 8 --warmup 3 --named cfg5 --skip-cpu > $O/${TAG}_bench_forced.json 2> $O/${TAG}_bench_forced.err
tail -2 $O/${TAG}_bench_forced.err | cut -c1-300; python -c "
import json; d=json.load(open('$O/${TAG}_bench_forced.json'))
for n in d['config']['named']: print(' named', n['name'], n['config']['envs_per_gpu'], 'envs/gpu', n['value']/1e6, 'M sims/s', n['ms_per_step'], 'ms e2e', n['e2e']['value']/1e6, n['config'].get('nccl'))"
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 2 --steps 6 --warmup 3 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err; cut -c1-160 $O/${TAG}_ref.json
