#!/bin/bash
# Bench every kernel-variant library (tools/build_variant.sh) plus the default one; one line each.
mkdir -p gpurun_out
for lib in default umeregrobust_b200/csrc/variants/*.so; do
  if [ "$lib" = default ]; then unset UME_LIB_PATH; else export UME_LIB_PATH=$PWD/$lib; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 "$@" 2>>gpurun_out/variants.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-60s value %8.0f  ms/step %.3f  moments %.3f ms/launch frac %.3f  stages %s' % ('$lib', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], {k: round(v['ms_per_step'],3) for k,v in d['stages'].items()}))"
done | tee -a gpurun_out/variants.log
