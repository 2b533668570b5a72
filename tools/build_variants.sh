#!/bin/bash
# tuning experiments: build the library once per set of -D flags into pgslam_b200/lib/var_<name>.so
# usage: tools/build_variants.sh name1 "-DX=1 -DY=2" name2 "..." ...
set -e
cd "$(dirname "$0")/.."
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  PGS_NVCC_EXTRA="$flags" python -m pgslam_b200.build --force > /dev/null
  cp pgslam_b200/lib/libpgslam_b200.so pgslam_b200/lib/var_$name.so
  echo "built var_$name.so with: $flags"
done
python -m pgslam_b200.build --force > /dev/null
