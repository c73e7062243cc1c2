#!/bin/bash
# usage (GPU box): bash scripts/e2e_sweep.sh "<bins> <threads> <gather> <payload>" ...   -- e2e leg only
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  timeout 500 python bench.py --steps 5 --no-cpu --no-text --e2e-bins $1 --e2e-threads $2 --gather-threads $3 --e2e-payload $4 > gpurun_out/e2e_sweep.json 2> gpurun_out/e2e_sweep.err || { echo "$cfg failed"; tail -3 gpurun_out/e2e_sweep.err; continue; }
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open("gpurun_out/e2e_sweep.json"))
e = d["e2e"]
print("bins/threads/gather/payload %-12s e2e %.2f ms  %.3e reads/s   single %.2f ms" % (sys.argv[1], e["ms_per_step"], e["value"], e["single_call"]["ms_per_step"]))
PY
done
