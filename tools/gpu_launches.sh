#!/bin/bash
# per-launch durations of one short bench run (ncu, serialised, cold caches: shares only)
TAG=${1:-l}; shift
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --full-reg-pairs 0 "$@" > gpurun_out/${TAG}_launches.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1e3 if u=="ns" else (v if u in ("us","usecond") else v*1e3 if u=="ms" else v)
    agg.setdefault(r[ki][:70],[]).append(v)
for k,v in agg.items(): print("%-72s n=%3d mean %9.1f us  last %9.1f us"%(k,len(v),sum(v)/len(v),v[-1]))
PY
