#!/bin/bash
# diagnostic build with fine stamps inside the fused cell phase
XG_EXTRA_NVCC_FLAGS="-DGK_FINE" python controllable_xgating_b200/build.py --force > /dev/null 2>&1
XG_PERSIST_TRACE=1 timeout 300 python scripts/greedy_once.py 4 2>&1 | grep "trace" | tail -8
