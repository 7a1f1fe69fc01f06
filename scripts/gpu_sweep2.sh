#!/bin/bash
for v in "XG_M2=9" "XG_M2=6" "XG_M2=7" "XG_M2=5" "XG_M1=4" "XG_M1=3" "XG_EARLY=0" "XG_EARLY=1" "XG_EARLY=2" "XG_M2=6 XG_M1=4"; do
  echo "== $v"; env $v timeout 300 python scripts/profile_path.py greedy 3 2>&1 | grep "decode_persistent"
done
