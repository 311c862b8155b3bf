#!/bin/bash
# Timing experiments for the v2 joint kernel: which role sets the pace?  (results are garbage with RNNTB200_DBG != 0)
export RNNTB200_TC_VARIANT=2
for d in 0 1 2 4 8 3 12 15; do
  echo "==== DBG=$d"
  RNNTB200_DBG=$d timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    print({k: v for k, v in j['kernels_ms'].items() if 'joint' in k}, 'step_ms', round(j['ms_per_step'], 2))
"
done
