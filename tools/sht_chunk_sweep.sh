#!/bin/bash
# Sweep of the L2-chunked transform (runs per chunk x streams) on one GPU: ms per step and per-group times of the bench workload.
# usage (GPU box): bash tools/sht_chunk_sweep.sh > gpurun_out/sht_sweep.log
for cfg in "0 1" "1 2" "1 3" "2 1" "2 2" "2 3" "2 4" "4 2" "4 3" "8 2"; do
  set -- $cfg
  echo "== XFB_SHT_CHUNK=$1 XFB_SHT_STREAMS=$2"
  XFB_SHT_CHUNK=$1 XFB_SHT_STREAMS=$2 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    ln = ln.strip()
    if ln.startswith('{'):
        d = json.loads(ln)
        g = d['roofline']['groups']
        print('ms_per_step %.3f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], ' '.join('%s=%.2f' % (k, v['ms_per_step']) for k, v in g.items()))
    elif ln: print(ln[:300])
"
done
