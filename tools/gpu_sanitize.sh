#!/bin/bash
# compute-sanitizer over the GPU parity tests of the kernels added / changed in round 2 (run under gpurun).
set -u
mkdir -p gpurun_out
P=gpurun_out/r02_sanitizer
: > $P.log
run() {  # tool, label, pytest args...
  tool=$1; label=$2; shift 2
  echo "== $tool: $label" | tee -a $P.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > ${P}_tmp.log 2>&1
  echo "rc=$?" | tee -a $P.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" ${P}_tmp.log | tail -4 | tee -a $P.log
}
run memcheck "training row: moment / distance / solve backward" tests/test_gpu_training.py
run memcheck "hypothesis scoring, kNN, spatial variance, voxel de-duplication" tests/test_gpu_corr.py
run memcheck "hot path vs reference golden, ball query, descriptors, rigid" tests/test_gpu_parity.py -k "hot_path or ball_query or ortho or rigid or knn1"
run memcheck "pair launch, Gumbel top-k, rotation error, result pack, graphs" tests/test_gpu_round2.py -k "pair_launch or gumbel or rotation_error or result_pack or graphs_of or presplit"
run racecheck "stable grid + warp moment kernel + distance kernels" tests/test_gpu_parity.py -k "hot_path or moments"
run racecheck "hypothesis scoring (shared-memory K-best columns), distance backward" tests/test_gpu_corr.py tests/test_gpu_training.py -k "golden or backward_kernels"
