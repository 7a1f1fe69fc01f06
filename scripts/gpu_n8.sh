#!/bin/bash
# 8-GPU call: config-4 bench (64 captions per GPU) with and without the all-reduce overlap
mkdir -p gpurun_out
for ov in 1 0; do
XG_DP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_ov$ov.json 2> gpurun_out/bench_n8_ov$ov.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n8_ov$ov.json')); print('overlap=$ov', d['value'], d['e2e']['value'], d['train']['value'], d['train']['ms_per_step'])"
done
