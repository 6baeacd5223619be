set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 2>gpurun_out/r2z_bench_bnn.err | tail -1 > gpurun_out/r2z_bench_bnn.json
for w in logreg svgd vae ar1; do python bench.py --workload $w --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2z_bench_$w.json; done
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2z_bench_reference.json
BRN_BENCH_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2z_launches_bnn.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-api > /dev/null 2>&1
BRN_BENCH_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"umma_nt|bnn_mid4|mf_stats|sample_w1_group" -s 12 -c 5 -o gpurun_out/r2z_ncu_full_bnn python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"umma_nt|svgd|EpiBernoulli" -s 30 -c 12 -o gpurun_out/r2z_ncu_full_svgd python bench.py --workload svgd --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -12
