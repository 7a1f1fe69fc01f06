#!/bin/bash
# end-of-round check: the GPU test suite, smoke(), the default bench line of both arms
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print('bench', d['value'], d['e2e']['value'], d['train']['value'], d['beam']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
