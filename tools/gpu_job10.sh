#!/bin/bash
TAG=${1:-job10}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
KGE_B200_FUSED_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_timing.json 2> gpurun_out/${TAG}_timing.err
grep -A1 "fused timing" gpurun_out/${TAG}_timing.err | tail -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fused' -s 6 -c 2 -f -o gpurun_out/${TAG}_fusedprof python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<P
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value %.1fM e2e %.1fM ms %.4f frac %.3f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['roofline']['frac']))
print({k:round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})
P
