#!/bin/bash
# tests + one bench line (no CPU legs)
set -u
mkdir -p gpurun_out
TAG=${1:-q}; shift
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --no-cpu-baseline --full-reg-pairs 0 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null || cat gpurun_out/${TAG}_bench.json
