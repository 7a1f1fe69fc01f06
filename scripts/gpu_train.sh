#!/bin/bash
mkdir -p gpurun_out
XG_PERSIST_TRACE=1 timeout 300 python scripts/profile_path.py train 3 > gpurun_out/prof_train_g.txt 2>&1; grep "trace" gpurun_out/prof_train_g.txt | tail -3; grep -v "trace\|Warn" gpurun_out/prof_train_g.txt | head -12
( timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
