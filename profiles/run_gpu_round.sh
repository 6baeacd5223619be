#!/bin/bash
# One gpurun call: GPU parity tests, bench lines per workload, ncu launch list and one --set full capture per top kernel.
# usage (from the repo root, under gpurun):  bash profiles/run_gpu_round.sh <tag> [quick]
TAG=${1:-r1x}
MODE=${2:-full}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.txt
tail -3 $O/${TAG}_pytest_gpu.txt
for wl in bnn logreg svgd vae ar1; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 > $O/${TAG}_bench_$wl.json 2> $O/${TAG}_bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_$wl.json")); print("$wl", round(d["ms_per_step"],4), "ms/step; e2e", round(d["e2e"]["ms_per_step"],4), round(d["roofline"].get("frac"),4), d["roofline"].get("whole_step",{}).get("frac"), "cpu", d["cpu_baseline"]["value"], "ours", d["value"])
except Exception as e: print("bad json", e); print(open("$O/${TAG}_bench_$wl.err").read()[-1500:])
PY
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>$O/${TAG}_bench_reference.err; tail -c 600 $O/${TAG}_bench_reference.json
[ "$MODE" = quick ] && exit 0
for wl in bnn logreg svgd vae; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_$wl.csv \
  python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline > $O/${TAG}_ncu_$wl.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'umma_nt|bnn_mid|sample_w1|mf_stats' -s 12 -c 6 -f -o $O/${TAG}_prof_bnn \
  python bench.py --workload bnn --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_full_bnn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'umma_nt' -s 30 -c 4 -f -o $O/${TAG}_prof_logreg \
  python bench.py --workload logreg --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_full_logreg.log 2>&1
ls -la $O | tail -30
