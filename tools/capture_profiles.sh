#!/bin/bash
# on the GPU box: the evidence bundle copied into profiles/ (launch list of the bench command,
# full ncu capture of the dominant kernel, bench line).  usage: tools/capture_profiles.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 4 -c 1 -o gpurun_out/${tag}_match_full \
    python tools/profile_run.py 96 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_kernel -c 1 -o gpurun_out/${tag}_knn10_full \
    python tools/profile_run.py 96 1 > /dev/null 2>&1
cat gpurun_out/${tag}_bench_n1.json
