#!/bin/bash
# One full ncu capture (with source counters) of the moment kernel inside a short bench run.
TAG=${1:-prof}; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:moments -s 2 -c 1 \
    -o gpurun_out/${TAG}_moments -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
