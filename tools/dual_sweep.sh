#!/bin/bash
# Sweep of the two-stream mode (SM shares of the two Jacobi launches) on one GPU, bench workload.  usage: bash tools/dual_sweep.sh [runs]
R=${1:-128}
for cfg in "0 50 20" "1 148 148" "1 74 30" "1 60 24" "1 50 20" "1 40 16" "1 50 30" "1 64 16"; do
  set -- $cfg
  echo "== XFB_DUAL=$1 BIG=$2 SMALL=$3 runs=$R"
  XFB_DUAL=$1 XFB_DUAL_BIG=$2 XFB_DUAL_SMALL=$3 python bench.py --steps 10 --warmup 3 --no-cpu --runs $R 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    ln = ln.strip()
    if ln.startswith('{'):
        d = json.loads(ln)
        print('ms_per_step %.3f  single-stream %.3f  value %.0f  e2e %.0f' % (d['ms_per_step'], d['ms_per_step_single_stream'], d['value'], d['e2e']['value']))
    elif ln: print(ln[:300])
"
done
