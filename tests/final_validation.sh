#!/bin/bash
# Round-end validation on the GPU box: parity suite, smoke, the three bench workloads, per-layer table and ncu launch list.
# usage (under gpurun): bash tests/final_validation.sh <tag>
tag=${1:-final}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --workload flow --steps 10 --warmup 3 > gpurun_out/${tag}_flow.json 2> gpurun_out/${tag}_flow.err
python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/${tag}_train.json 2> gpurun_out/${tag}_train.err
SHINEON_RAW=1 python tests/profile_layers.py 32 fp16x3 > gpurun_out/${tag}_layers.txt 2>&1
SHINEON_RAW=1 SHINEON_CUPROF=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${tag}_launches.csv python tests/profile_layers.py 32 fp16x3 > /dev/null 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json")); r = d["roofline"]
print("tryon", d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["conv_share_of_step"], d["cpu_baseline"]["value"], d["clocks"])
d = json.load(open("gpurun_out/${tag}_flow.json"))
print("flow", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gather_ops"]["resample2d"]["frac_of_hbm_peak"], d["gather_ops"]["correlation"]["frac_of_hbm_peak"])
d = json.load(open("gpurun_out/${tag}_train.json"))
print("train", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
