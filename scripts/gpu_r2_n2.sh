#!/bin/bash
# Round-2 two-GPU visit: bench.py under torchrun with the named configs of N = 2 (configs[2] strong-scaled) plus, forced, the
# configs[4] step with its NCCL legs (cross-rank replay sample + gradient-mean all-reduce) and configs[3].
TAG=${1:-r2n2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 12 --warmup 4 > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err
tail -3 $O/${TAG}_bench_default.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_default.json')); print('N=2 default:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'e2e', d['e2e']['value']/1e6)
for n in d['config']['named']: print(' named', n['name'], n['scaling'], n['config']['envs_per_gpu'], 'envs/gpu', n['value']/1e6, 'M sims/s', n['ms_per_step'], 'ms e2e', n['e2e']['value']/1e6, 'roof', n['roofline']['frac'], 'reroot', n['roofline']['reroot']['frac'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 8 --warmup 4 --named cfg5,cfg4 --skip-cpu > $O/${TAG}_bench_forced.json 2> $O/${TAG}_bench_forced.err
tail -3 $O/${TAG}_bench_forced.err; python -c "
import json; d=json.load(open('$O/${TAG}_bench_forced.json')); print('N=2 forced:', d['value']/1e6, 'M sims/s')
for n in d['config']['named']: print(' named', n['name'], n['config']['envs_per_gpu'], 'envs/gpu', n['value']/1e6, 'M sims/s', n['ms_per_step'], 'ms e2e', n['e2e']['value']/1e6, n['config'].get('nccl'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 6 --warmup 3 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cut -c1-300 $O/${TAG}_bench_ref.json
