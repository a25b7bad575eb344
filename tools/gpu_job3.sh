#!/bin/bash
TAG=${1:-job3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/${TAG}_fused.log 2>&1; echo "fused rc=$?" | tee gpurun_out/${TAG}_rc.txt
KGE_B200_SPLIT_TRUNC=1 timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/${TAG}_trunc.log 2>&1; echo "trunc rc=$?" | tee -a gpurun_out/${TAG}_rc.txt
KGE_B200_FUSED_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_timing.json 2> gpurun_out/${TAG}_timing.err
KGE_B200_FUSED_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --batch 1000 > gpurun_out/${TAG}_timing_B1000.json 2> gpurun_out/${TAG}_timing_B1000.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_fused.log; tail -3 gpurun_out/${TAG}_trunc.log; grep "fused timing" gpurun_out/${TAG}_timing.err | tail -2; grep "fused timing" gpurun_out/${TAG}_timing_B1000.err | tail -2
python - <<P
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value %.1fM e2e %.1fM ms %.4f frac %.3f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['roofline']['frac']))
print({k:round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})
P
