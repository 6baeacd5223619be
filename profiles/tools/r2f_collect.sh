# final round-2 collection on ONE GPU: bench lines of every workload, launch lists, ncu full captures of the dominant kernels
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 2>gpurun_out/r2f_bench_bnn.err | tail -1 > gpurun_out/r2f_bench_bnn.json
for w in logreg svgd vae ar1; do python bench.py --workload $w --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2f_bench_$w.json; done
BRN_BENCH_NO_PREPARED=1 python bench.py --workload logreg --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2f_bench_logreg_unprepared.json
BRN_LINEAR_FLASH=0 python bench.py --workload logreg --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2f_bench_logreg_staged.json
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2f_bench_reference.json
BRN_BENCH_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2f_launches_bnn.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-api > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r2f_launches_logreg.csv python bench.py --workload logreg --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2f_launches_svgd.csv python bench.py --workload svgd --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# full captures are summarised ON THE BOX (the .ncu-rep files together exceed what gpurun copies back); only the K2 report, which
# carries the source page of the new kernel, travels
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:linear_flash_kernel -s 2 -c 1 -o gpurun_out/r2f_ncu_full_logreg python bench.py --workload logreg --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python profiles/ncu_summary.py gpurun_out/r2f_ncu_full_logreg.ncu-rep > gpurun_out/r2f_ncu_full_logreg_summary.txt
timeout -s KILL 300 ncu --set full --clock-control none -k regex:"linear_flash_kernel|umma_nt|svgd_select" -s 8 -c 10 -o /tmp/r2f_ncu_full_svgd python bench.py --workload svgd --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/r2f_ncu_full_svgd.ncu-rep > gpurun_out/r2f_ncu_full_svgd_summary.txt
timeout -s KILL 400 ncu --set full --clock-control none -k regex:"umma_nt" -s 30 -c 16 -o /tmp/r2f_ncu_full_vae python bench.py --workload vae --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/r2f_ncu_full_vae.ncu-rep > gpurun_out/r2f_ncu_full_vae_summary.txt
BRN_BENCH_NO_GRAPH=1 timeout -s KILL 300 ncu --set full --clock-control none -k regex:"umma_nt|bnn_mid4|mf_stats|sample_w1_group" -s 12 -c 5 -o /tmp/r2f_ncu_full_bnn python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api > /dev/null 2>&1
python profiles/ncu_summary.py /tmp/r2f_ncu_full_bnn.ncu-rep > gpurun_out/r2f_ncu_full_bnn_summary.txt
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2f_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.txt 2>&1
du -sh gpurun_out
ls -la gpurun_out | tail -30
