#!/bin/bash
for v in "XG_TM1=5" "XG_TM1=4" "XG_TM1=3"; do echo "== $v"; env $v XG_PERSIST_TRACE=1 timeout 300 python scripts/profile_path.py train 3 2>&1 | grep "train trace\|train_decode_persistent" | tail -4; done
timeout 900 python -m pytest tests -m gpu -q -x -k "train or persistent or golden" 2>&1 | tail -2
