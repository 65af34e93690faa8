#!/bin/bash
for lib in default umeregrobust_b200/csrc/variants/*.so; do
  if [ "$lib" = default ]; then unset UME_LIB_PATH; else export UME_LIB_PATH=$PWD/$lib; fi
  echo "$lib: $(python tools/bench_corr.py 1024 10000 2>&1 | tail -1)"
done
