#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_persistent.py tests/test_gpu_parity.py -m gpu -q -x -k "matches_unfused_and_golden or full_size_with_eos or partial_batches or greedy" ) > gpurun_out/pytest_grouped.log 2>&1
tail -5 gpurun_out/pytest_grouped.log
XG_PERSIST_TRACE=1 timeout 300 python scripts/greedy_once.py 4 2>&1 | grep "trace" | grep -v "step 3" | tail -5
for v in "XG_NO_GROUPED=0" "XG_NO_GROUPED=1"; do
env $v timeout 600 python bench.py --steps 10 --warmup 3 --skip-extra --skip-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print('$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'])"
done
