#!/bin/bash
# Build a kernel-variant copy of the library: tools/build_variant.sh NAME -DUME_MOMENTS_MINB=4 ...
# -> umeregrobust_b200/csrc/variants/libumereg_NAME.so   (use with UME_LIB_PATH=...)
set -e
NAME=$1; shift
CS=umeregrobust_b200/csrc
mkdir -p $CS/variants/obj_$NAME
for f in $CS/*.cu; do
  o=$CS/variants/obj_$NAME/$(basename ${f%.cu}).o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include "$@" -c $f -o $o &
done
wait
nvcc -shared -o $CS/variants/libumereg_$NAME.so $CS/variants/obj_$NAME/*.o
rm -rf $CS/variants/obj_$NAME
echo built $CS/variants/libumereg_$NAME.so
