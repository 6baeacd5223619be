#!/bin/bash
# One gpurun --gpus N call: every workload's bench line at N ranks (torchrun, NCCL), as the driver launches it.
# usage: bash profiles/run_gpu_scale.sh <tag> <N> [workloads...]
TAG=${1:-r1x}
N=${2:-2}
shift 2
WLS=${@:-bnn logreg svgd vae}
O=gpurun_out
mkdir -p $O
for wl in $WLS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $wl --steps 50 --warmup 5 > $O/${TAG}_scale${N}_$wl.json 2> $O/${TAG}_scale${N}_$wl.err
  echo "rc=$? $wl"; tail -c 1200 $O/${TAG}_scale${N}_$wl.json; tail -3 $O/${TAG}_scale${N}_$wl.err
done
