#!/bin/bash
# Round-2 GPU battery (run under gpurun): parity tests, default bench line, reference arm, small stream bench.
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
lscpu | head -20 > gpurun_out/${TAG}_lscpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/${TAG}_bench_ref.json
timeout 600 python bench.py --workload tiny_stream --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_tiny_stream.json 2> gpurun_out/${TAG}_bench_tiny_stream.err; echo "tiny_stream rc=$?"
cat gpurun_out/${TAG}_bench_tiny_stream.json; tail -5 gpurun_out/${TAG}_bench_tiny_stream.err
