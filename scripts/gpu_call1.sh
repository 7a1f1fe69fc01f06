#!/bin/bash
# round 2, call 1: cluster probe, full GPU test-suite, smoke, short bench
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cluster_probe scripts/microbench/cluster_probe.cu && timeout 120 /tmp/cluster_probe > gpurun_out/cluster_probe.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 3000 gpurun_out/bench_a.json
cat gpurun_out/cluster_probe.txt
