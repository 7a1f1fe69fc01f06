#!/bin/bash
mkdir -p gpurun_out
XG_PERSIST_TRACE=1 timeout 300 python scripts/profile_path.py train 1 2>&1 | grep "bwd trace" | tail -4
timeout 300 python scripts/profile_path.py train 3 > gpurun_out/prof_train_g.txt 2>&1; grep -v "trace\|Warn" gpurun_out/prof_train_g.txt | head -8
( timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
