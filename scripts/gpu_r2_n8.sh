#!/bin/bash
# Round-2 eight-GPU visit: the driver's own command line for N = 8 (native arm, then the reference arm).
TAG=${1:-r2n8}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 12 --warmup 4 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -3 $O/${TAG}_bench.err | cut -c1-300; python -c "
import json; d=json.load(open('$O/${TAG}_bench.json')); print('N=8:', d['value']/1e6, 'M sims/s', d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['clocks'])
for n in d['config']['named']: print(' named', n['name'], n['scaling'], n['config']['envs_per_gpu'], 'envs/gpu', n['value']/1e6, 'M sims/s', n['ms_per_step'], 'ms e2e', n['e2e']['value']/1e6, 'roof', n['roofline']['frac'], 'reroot', n['roofline']['reroot']['frac'], n['config'].get('nccl'))"
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 6 --warmup 3 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err; cut -c1-250 $O/${TAG}_ref.json
