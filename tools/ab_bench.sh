#!/bin/bash
# A/B two builds of the library on ONE box (boxes of the pool differ by a few percent):
#   cp ultrasphere_b200/lib/libultrasphere_b200.so ab_libs/old.so ; <edit, make> ; cp ... ab_libs/new.so
#   gpurun -- 'US_ONLY_SPEC=hopf bash tools/ab_bench.sh'
# Alternates old/new twice through tools/config_bench.py (ULTRASPHERE_B200_LIB selects the build).
for i in 1 2; do for v in old new; do
  ULTRASPHERE_B200_LIB=$PWD/ab_libs/$v.so US_ONLY=${US_ONLY_SPEC:-hopf} timeout 300 python tools/config_bench.py 2>/dev/null | python -c "
import json,sys
c=json.load(sys.stdin)
for k,v in c.items():
    if isinstance(v,dict): print('$v', k, round(v['from_ms'],4), round(v['to_ms'],4))"
done; done
