#!/bin/bash
# Mnemonic counts of the built objects (profiles/sass_tcgen05.txt): run after _lib.build().
B=umeregrobust_b200/csrc/build
for o in cdist_tc moments corr rigid cdist_bwd ortho grid; do
  echo; echo "== $o.o"; cuobjdump -sass $B/$o.o > /tmp/$o.sass
  for m in UTCHMMA LDTM UTMALDG UTCBAR SYNCS FFMA2 FADD2 "LDG.E.128" "REDG.E.ADD.F32x4" REDUX ATOMS "LDS.128"; do
    c=$(grep -c -- "$m" /tmp/$o.sass); [ "$c" != 0 ] && echo "$m $c"
  done
done
