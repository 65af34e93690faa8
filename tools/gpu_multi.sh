#!/bin/bash
# Multi-GPU bench lines under `gpurun --gpus N`: tools/gpu_multi.sh TAG N workload[:extra args] ...
# Outputs gpurun_out/TAG_nN_<workload>.json (one JSON line each) + .err
set -u
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${TAG}_n${N}_smi.txt 2>&1
nproc > gpurun_out/${TAG}_n${N}_nproc.txt
port=29511
for spec in "$@"; do
  wl=${spec%%:*}; extra=""; [[ "$spec" == *:* ]] && extra=${spec#*:}
  port=$((port+1))
  out=gpurun_out/${TAG}_n${N}_${wl}
  t0=$(date +%s)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --workload $wl --no-cpu-baseline $extra > $out.json 2> $out.err
  echo "$wl rc=$? $(( $(date +%s) - t0 )) s"
  python tools/show_bench.py $out.json || tail -5 $out.err
done
