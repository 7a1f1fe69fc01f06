#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_persistent.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/pytest_grouped.log 2>&1
tail -5 gpurun_out/pytest_grouped.log
for v in "XG_NO_GROUPED=0" "XG_NO_GROUPED=1"; do
env $v timeout 600 python bench.py --steps 10 --warmup 3 --skip-extra --skip-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print('$v', d['value'], d['ms_per_step'], d['e2e']['value'], [(k['name'], round(k['us_per_launch'],1)) for k in d['roofline']['top5']])"
done
