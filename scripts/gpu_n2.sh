#!/bin/bash
# 2-GPU call: hardware DP parity test + config-4 bench with and without the all-reduce overlap
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_dp.py -m gpu -q -s ) > gpurun_out/r2_dp_parity_n2.txt 2>&1
tail -6 gpurun_out/r2_dp_parity_n2.txt
for ov in 1 0; do
XG_DP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_ov$ov.json 2> gpurun_out/bench_n2_ov$ov.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_ov$ov.json')); print('overlap=$ov', d['value'], d['train']['value'], d['train']['ms_per_step'])"
done
