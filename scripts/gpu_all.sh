#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for w in greedy train beam; do timeout 300 python scripts/profile_path.py $w 3 2>&1 | grep -v Warn | head -5; done
