#!/bin/bash
# Round-end measurement set (one GPU): bench lines of every workload, the CPU reference arm, the
# correlator micro-benchmark.  Outputs: gpurun_out/final_*.json
mkdir -p gpurun_out
for wl in kitti_b64_n1024_c32 nuscenes_b64_n1024_c32 rotkitti_b32_n2048_c64 kitti_b1_n512_c32; do
  extra=""; [ $wl != kitti_b64_n1024_c32 ] && extra="--no-cpu-baseline"
  timeout 900 python bench.py --workload $wl $extra > gpurun_out/final_$wl.json 2> gpurun_out/final_$wl.err
  python tools/show_bench.py gpurun_out/final_$wl.json
done
timeout 900 python bench.py --cta-moments 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/final_cta_kernel.json 2>/dev/null
python tools/show_bench.py gpurun_out/final_cta_kernel.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference_arm.json 2> gpurun_out/final_reference_arm.err
cat gpurun_out/final_reference_arm.json | cut -c1-600
timeout 600 python tools/bench_corr.py > gpurun_out/final_corr.log 2>&1; tail -5 gpurun_out/final_corr.log
