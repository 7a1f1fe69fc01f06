#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python -c "
import json; d=json.load(open('gpurun_out/bench_full.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['train']['value'], d['beam']['value'], d['cpu_baseline'])"
