#!/bin/bash
mkdir -p gpurun_out
KGE_B200_FUSED_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2g_timing.json 2> gpurun_out/r2g_timing.err
grep -A1 "fused timing" gpurun_out/r2g_timing.err | tail -4
echo HALFLOAD
KGE_B200_FUSED_HALFLOAD=1 KGE_B200_FUSED_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2g_timing_half.json 2> gpurun_out/r2g_timing_half.err
grep -A1 "fused timing" gpurun_out/r2g_timing_half.err | tail -4
timeout 600 python -m pytest tests/test_dist.py tests/test_gpu_fused.py -q -m gpu > gpurun_out/r2g_tests.log 2>&1; tail -3 gpurun_out/r2g_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<P
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print('value %.1fM e2e %.1fM ms %.4f frac %.3f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['roofline']['frac']))
print({k:round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})
P
