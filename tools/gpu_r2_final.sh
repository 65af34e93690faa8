#!/bin/bash
# Round-2 measurement set on ONE GPU (run under gpurun): parity tests, the bench line of every workload, the CPU
# reference arm, the launch list and full ncu captures of the hot kernels.  Outputs: gpurun_out/r02_*
set -u
mkdir -p gpurun_out
P=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > ${P}_smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
timeout 900 python bench.py > ${P}_bench_kitti_b64_n1024_c32.json 2> ${P}_bench.err; echo "bench rc=$?"
python tools/show_bench.py ${P}_bench_kitti_b64_n1024_c32.json
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > ${P}_bench_reference_arm.json 2> ${P}_bench_ref.err; echo "ref rc=$?"
cut -c1-400 ${P}_bench_reference_arm.json
for wl in kitti_b1_n512_c32 nuscenes_b64_n1024_c32 rotkitti_b32_n2048_c64 tiny_stream; do
  timeout 900 python bench.py --workload $wl --no-cpu-baseline > ${P}_bench_$wl.json 2> ${P}_bench_$wl.err; echo "$wl rc=$?"
  python tools/show_bench.py ${P}_bench_$wl.json
done
timeout 900 python bench.py --cta-moments 1 --no-cpu-baseline --e2e-steps 0 --full-reg-pairs 0 > ${P}_bench_kitti_b64_cta_kernel.json 2>/dev/null
python tools/show_bench.py ${P}_bench_kitti_b64_cta_kernel.json
timeout 600 python tools/bench_corr.py 2500 10000 2>&1 | tail -1 | tee ${P}_corr.log
timeout 600 python tools/bench_corr.py 1024 10000 2>&1 | tail -1 | tee -a ${P}_corr.log
# launch list (serialised, cold caches: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --full-reg-pairs 0 > ${P}_launches.log 2>&1
[ "${1:-}" = nocapture ] && exit 0
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --full-reg-pairs 0"
for spec in moments:moments_warp_kernel:2 cdist:cdist_tc_kernel:2 gridrank:grid_rank_kernel:2; do
  IFS=: read tag kern skip <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -o ${P}_${tag} -f $BENCH > ${P}_${tag}_ncu.log 2>&1
  ls -la ${P}_${tag}.ncu-rep
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:corr_score_kernel -s 2 -c 1 -o ${P}_corr -f \
    python tools/bench_corr.py 256 10000 > ${P}_corr_ncu.log 2>&1
ls -la ${P}_corr.ncu-rep
