#!/bin/bash
# quick GPU iteration: parity tests + bench lines (A/B via env) + optional ncu --set full of named kernels
# usage: bash profiles/run_gpu_quick.sh <tag> "<workloads>" ["<ncu kernel regex>" [skip count]]
TAG=${1:-r1x}; WLS=${2:-bnn}; KRE=$3; SKIP=${4:-12}; CNT=${5:-4}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.txt
tail -3 $O/${TAG}_pytest_gpu.txt
for wl in $WLS; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 > $O/${TAG}_bench_$wl.json 2> $O/${TAG}_bench_$wl.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_$wl.json")); print("$wl", round(d["ms_per_step"],4), "ms/step; e2e", round(d["e2e"]["ms_per_step"],4), d["roofline"].get("frac"), d["roofline"].get("whole_step"), d["roofline"]["stage_share_of_step"])
except Exception as e: print("bad json", e); print(open("$O/${TAG}_bench_$wl.err").read()[-1500:])
PY
done
if [ -n "$KRE" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -f -o $O/${TAG}_prof \
  python bench.py --workload bnn --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
tail -2 $O/${TAG}_ncu_full.log
fi
