#!/bin/bash
# Timing experiments for the joint kernel (RNNTB200_TC_VARIANT, default 3): which role sets the pace?
# RNNTB200_DBG bits: 1 no epilogue math, 2 no tanh (v2), 4 no W TMA / waits, 8 no producers, 64 extra per-stage fence,
# 128 no acc_empty wait, 256 no tcgen05.ld in the epilogue.  Results are garbage with RNNTB200_DBG != 0 -- timing only.
export RNNTB200_TC_VARIANT=${RNNTB200_TC_VARIANT:-3}
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
