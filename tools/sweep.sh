#!/bin/bash
# Tuning sweep for the persistent wide kernel (run under gpurun). Usage: tools/sweep.sh "VAR=val VAR=val" ...
export ZYG_BENCH_CACHE=/tmp/zyg_cache
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())
print('total %.0f Mrays/s e2e %.0f' % (r['value'], r['e2e']['value']))
for k,v in r['classes'].items(): print('  %-20s %7.0f Mrays/s %6.2f ms nodes %.2f tris %.2f stack %d' % (k, v['mrays_s'], v['ms'], v['nodes_per_ray'], v['tris_per_ray'], v['max_stack']))
"
done
