#!/bin/bash
# usage: bash scripts/quick.sh <tag>   -- GPU box: parity tests + kernel-only bench
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.3e reads/s  ms/step %.3f  kernels %s  stages %s" % (d["value"], d["ms_per_step"], d["roofline"]["other_kernel"], d["stages_ms"]))
PY
