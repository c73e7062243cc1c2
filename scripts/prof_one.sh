#!/bin/bash
# usage: bash scripts/prof_one.sh <tag> <kernel-regex>   -- GPU box: one ncu --set full capture of one kernel
TAG=${1:-x}; K=${2:-k_discover_tiles}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
