#!/bin/bash
mkdir -p gpurun_out
XG_PERSIST_TRACE=1 timeout 300 python scripts/profile_path.py beam 1 2>&1 | grep "trace" | tail -6 > gpurun_out/prof_beam_g.txt; cat gpurun_out/prof_beam_g.txt
timeout 300 python scripts/profile_path.py beam 3 2>&1 | grep -v Warn | head -4
( timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
