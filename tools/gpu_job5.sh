#!/bin/bash
TAG=${1:-job5}
mkdir -p gpurun_out
KGE_B200_FUSED_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_timing.json 2> gpurun_out/${TAG}_timing.err
grep -A1 "fused timing" gpurun_out/${TAG}_timing.err | tail -4
timeout 900 python -m pytest tests/test_dist.py -q -m gpu -x > gpurun_out/${TAG}_dist.log 2>&1; echo "dist rc=$?"; tail -15 gpurun_out/${TAG}_dist.log
