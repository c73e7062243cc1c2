#!/bin/bash
# The prepared kernel experiments (k_discover_tiles candidate queue /
# aggregated slots, k_call_observe locate-and-prefetch), end to end.
#   here (no GPU):  bash scripts/try_disc_variants.sh build
#   GPU box:        bash scripts/try_disc_variants.sh run        (parity tests per variant, then the A/B bench)
# A variant is only worth adopting if every parity test passes with AVO_LIB pointing at it.
set -u
VARIANTS=(cq agg cqagg loc)
FLAGS=("-DAVO_DISC_CQ" "-DAVO_DISC_AGG" "-DAVO_DISC_CQ -DAVO_DISC_AGG" "-DAVO_OBS_LOCATE")
if [ "${1:-}" = "build" ]; then
  args=()
  for i in "${!VARIANTS[@]}"; do args+=("${VARIANTS[$i]}" "${FLAGS[$i]}"); done
  bash scripts/variants.sh "${args[@]}"
  ls -la avocado_b200/variants/
  exit 0
fi
mkdir -p gpurun_out
for v in "${VARIANTS[@]}"; do
  lib=$PWD/avocado_b200/variants/lib_$v.so
  [ -f "$lib" ] || { echo "$v: not built (run 'build' first)"; continue; }
  AVO_LIB=$lib timeout 600 python -m pytest tests -m gpu -x -q -k "parity or edges or allsites or cohort" > gpurun_out/pytest_var_$v.log 2>&1
  echo "$v: pytest rc=$? $(tail -1 gpurun_out/pytest_var_$v.log)"
done
bash scripts/variants.sh --run "${VARIANTS[@]}"
