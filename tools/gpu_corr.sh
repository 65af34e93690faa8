#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-c}
timeout 1200 python -m pytest tests/test_gpu_corr.py tests/test_gpu_round2.py -m gpu -x -q -k "corr or select or end_to_end or spatial" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/bench_corr.py 2500 10000 2>&1 | tail -4 | tee gpurun_out/${TAG}_corr.log
timeout 600 python tools/bench_corr.py 1024 10000 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_corr.log
