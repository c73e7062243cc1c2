#!/bin/bash
# usage (CPU box): bash scripts/variants.sh name1 "-DAVO_X=.." name2 "..."  -> avocado_b200/variants/lib_<name>.so
# usage (GPU box): bash scripts/variants.sh --run name1 name2 ...           -> kernel-only bench per variant
if [ "$1" = "--run" ]; then
  shift
  mkdir -p gpurun_out
  for v in default "$@"; do
    lib=avocado_b200/variants/lib_$v.so; [ "$v" = default ] && lib=avocado_b200/libavocado_b200.so
    AVO_LIB=$PWD/$lib timeout 600 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$v.json"))
    print("%-14s value %.3e  ms/step %.3f  %s  stages %s" % ("$v", d["value"], d["ms_per_step"], d["roofline"]["other_kernel"], d["stages_ms"]))
except Exception as e: print("$v failed", e)
PY
  done
  exit 0
fi
mkdir -p avocado_b200/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  make -C avocado_b200/csrc -j8 BUILD=build_$name TARGET=../variants/lib_$name.so EXTRA="$flags" > /dev/null 2>&1 || echo "build $name failed"
  grep -h "Used" avocado_b200/csrc/build_$name/avo_discover.ptxas.log avocado_b200/csrc/build_$name/avo_call.ptxas.log | grep -v "Used [0-9] reg\|Used 1[0-9] reg" | head -4
done
