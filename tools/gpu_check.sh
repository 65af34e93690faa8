#!/bin/bash
# Standard GPU battery (run under gpurun): parity tests, bench line, launch list, one full ncu capture
# of the moment kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-run}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:moments -s 2 -c 1 \
    -o gpurun_out/${TAG}_moments -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${TAG}_ncu_moments.log 2>&1
ls -la gpurun_out
