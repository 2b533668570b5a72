#!/bin/bash
# on the GPU box: per-kernel durations (ncu, one metric) of a profile_run batch for every variant build
# usage: tools/ncu_variants.sh <kernel-regex> [pairs]
cd "$(dirname "$0")/.."
cp pgslam_b200/lib/libpgslam_b200.so /tmp/default.so
for v in pgslam_b200/lib/var_*.so; do
  cp $v pgslam_b200/lib/libpgslam_b200.so
  echo "== $v"
  ncu -k "regex:$1" --metrics gpu__time_duration.sum --clock-control none --csv python tools/profile_run.py ${2:-96} 1 2>/dev/null | python -c "
import csv,sys,collections
lines=[l for l in sys.stdin if l.startswith('\"')]
agg=collections.OrderedDict()
for r in csv.DictReader(lines):
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    k=r['Kernel Name'].split('(')[0][-40:]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,v) in agg.items(): print('  %-42s x%-3d %9.0f us' % (k,c,v))"
done
cp /tmp/default.so pgslam_b200/lib/libpgslam_b200.so
