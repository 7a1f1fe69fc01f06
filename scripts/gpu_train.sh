#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_path.py train 3 > gpurun_out/prof_train_g.txt 2>&1; grep -v "trace\|Warn" gpurun_out/prof_train_g.txt | head -14
( timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
