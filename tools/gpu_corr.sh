#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-c}
timeout 1200 python -m pytest tests/test_gpu_corr.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/bench_corr.py 2500 10000 2>&1 | tail -2 | tee gpurun_out/${TAG}_corr.log
timeout 600 python tools/bench_corr.py 1024 10000 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_corr.log
