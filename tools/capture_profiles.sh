#!/bin/bash
# on the GPU box: the evidence bundle copied into profiles/ - one commit, one bundle.
#   launch list of a short bench command, launch list of one 96-pair batch on one stream, full ncu
#   captures of the five kernels that carry the step, the bench line itself.
# usage: tools/capture_profiles.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
    python bench.py --pool 192 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --parity-pairs 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_C2_96pairs_onestream.csv \
    python tools/profile_run.py 96 1 > /dev/null 2>&1
for k in match_kernel:5 knn_kernel:0 accumulate_kernel:5 select_pass_kernel:15 kd_local_kernel:0 normals_kernel:0; do
  name=${k%%:*}; skip=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:^${name} -s $skip -c 1 -o gpurun_out/${tag}_${name}_full \
      python tools/profile_run.py 96 1 > /dev/null 2>&1
done
cat gpurun_out/${tag}_bench_n1.json
