#!/bin/bash
# A/B of environment switches on the GPU box:  bash tests/ab_env.sh <tag> "<workloads>|<ENV=.. ENV=..>" ...
# e.g. "tryon,flow|SHINEON_LANES=1".  Each set runs the named bench workloads (no CPU baseline) and prints value / ms_per_step.
tag=$1; shift
i=0
for spec in "$@"; do
  i=$((i+1))
  wls=${spec%%|*}; envs=${spec#*|}
  for wl in ${wls//,/ }; do
    env $envs python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_${i}_${wl}.json 2> gpurun_out/${tag}_${i}_${wl}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${i}_${wl}.json"))
    print("[$envs] $wl value %.1f ms %.3f e2e %.1f frac %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
except Exception as e:
    print("[$envs] $wl FAILED", e, open("gpurun_out/${tag}_${i}_${wl}.err").read()[-800:])
PY
  done
done
