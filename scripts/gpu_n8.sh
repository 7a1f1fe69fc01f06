#!/bin/bash
# 8-GPU call: the bench line at N = 8 (greedy decode weak scaling + config 4: global batch 512, 64 captions per GPU)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json')); print('n8', d['value'], d['e2e']['value'], d['train']['value'], d['train']['ms_per_step'])"
