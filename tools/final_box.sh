#!/bin/bash
# Runs ON the GPU box (via gpurun): the evidence set of a round -- bench lines of every configuration, the ncu launch list of the bench command, and the
# `ncu --set full` capture condensed to summaries / per-line shares / DRAM traffic.  usage: tools/final_box.sh <tag>     (outputs: gpurun_out/<tag>_*)
tag=${1:-r02}
mkdir -p gpurun_out
for c in 2 3 4 5; do
  timeout 300 python bench.py --config $c --steps 24 --warmup 3 --repeats 3 --no-cpu-baseline > gpurun_out/${tag}_bench_config${c}_n1.json 2> gpurun_out/${tag}_bench_config${c}_n1.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
SKIP=16 bash tools/prof_box.sh $tag 16
# the full default run (with the CPU baseline leg) last: it is the line the driver will reproduce
timeout 600 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
