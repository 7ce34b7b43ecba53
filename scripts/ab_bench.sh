#!/bin/bash
# A/B the library variants under sugar_b200/lib/variants: prints per-kernel ms for each
for lib in sugar_b200/lib/libsugar_b200.so sugar_b200/lib/variants/*.so; do
  [ -f "$lib" ] || continue
  case "$lib" in *lib_stats*.so) continue;; esac
  echo "=== $lib"
  SGR_LIB_PATH=$PWD/$lib timeout 200 python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('views/s %.1f  ms/step %.3f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print('  '.join('%s=%.3f' % (k, v['ms']) for k, v in d['stages'].items()))"
done
