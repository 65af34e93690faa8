#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-q}; shift
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --no-cpu-baseline --full-reg-pairs 0 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench.json
bash tools/gpu_launches.sh ${TAG}
