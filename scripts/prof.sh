#!/bin/bash
# usage: bash scripts/prof.sh <tag> [full]   -- GPU box: parity tests, bench, ncu launch list (+ full capture)
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
if [ "$2" = "full" ]; then
  $CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
  $CMD > gpurun_out/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_call_join|k_call_observe|k_call_reduce|k_discover_tiles' -s 4 -c 4 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
