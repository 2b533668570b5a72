#!/bin/bash
# on the GPU box: bench every pgslam_b200/lib/var_*.so (and the default build, last)
cd "$(dirname "$0")/.."
cp pgslam_b200/lib/libpgslam_b200.so /tmp/default.so
for v in pgslam_b200/lib/var_*.so; do
  cp $v pgslam_b200/lib/libpgslam_b200.so
  echo "== $v"
  python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f e2e %.0f ok %d iters %.3f stage %s' % (d['value'], d['e2e']['value'], d['pairs_ok'], d['iterations_mean'], d['roofline']['stage_ms_last_step']))"
done
cp /tmp/default.so pgslam_b200/lib/libpgslam_b200.so
