#!/bin/bash
# One full ncu capture (with source counters) of one kernel: tools/gpu_ncu_kernel.sh TAG KERNEL_REGEX SKIP -- command...
TAG=$1; KERN=$2; SKIP=$3; shift 4
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KERN -s $SKIP -c 1 \
    -o gpurun_out/${TAG} -f "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}.ncu-rep
