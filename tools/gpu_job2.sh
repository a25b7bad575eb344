#!/bin/bash
TAG=${1:-job2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -q -m gpu > gpurun_out/${TAG}_fused.log 2>&1; echo "fused rc=$?" | tee gpurun_out/${TAG}_rc.txt
KGE_B200_FUSED_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_timing.json 2> gpurun_out/${TAG}_timing.err
KGE_B200_FUSED_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --batch 1000 > gpurun_out/${TAG}_timing_B1000.json 2> gpurun_out/${TAG}_timing_B1000.err
tail -5 gpurun_out/${TAG}_fused.log; grep "fused timing" gpurun_out/${TAG}_timing.err | tail -4; grep "fused timing" gpurun_out/${TAG}_timing_B1000.err | tail -4
