#!/bin/bash
# A/B of env knobs: bash profiles/run_gpu_knobs.sh <tag> <workload> <ENV_NAME> "<values>"
TAG=$1; WL=$2; KNOB=$3; VALS=$4
O=gpurun_out; mkdir -p $O
for v in $VALS; do
  env $KNOB=$v timeout 200 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_${WL}_${KNOB}_$v.json 2> $O/${TAG}_${WL}_${KNOB}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_${WL}_${KNOB}_$v.json")); print("$WL $KNOB=$v", round(d["ms_per_step"],4), "ms/step", d["roofline"].get("frac"), d["roofline"]["stage_share_of_step"])
except Exception as e: print("bad json", e); print(open("$O/${TAG}_${WL}_${KNOB}_$v.err").read()[-800:])
PY
done
