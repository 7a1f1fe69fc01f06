#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_dp.py -m gpu -q -s -k "64" ) > gpurun_out/r2_dp_parity_n2b.txt 2>&1
grep -v "^$" gpurun_out/r2_dp_parity_n2b.txt | grep -i "error\|assert\|Traceback\|File\|worst" | head -40
