#!/bin/bash
# A/B timing of variants of the grouped greedy kernel on ONE box (boxes differ by several percent)
for v in "XG_L2_HINT=1" "XG_L2_HINT=0" "XG_L2_HINT=1" "XG_L2_HINT=0"; do
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --skip-extra --skip-cpu > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print('$v', d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'])"
done
