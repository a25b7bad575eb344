#!/bin/bash
TAG=${1:-job4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_umma.py -q -m gpu -x > gpurun_out/${TAG}_fused.log 2>&1; echo "fused rc=$?" | tee gpurun_out/${TAG}_rc.txt
KGE_B200_FUSED_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_timing.json 2> gpurun_out/${TAG}_timing.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 29600 > gpurun_out/${TAG}_bench_B29600.json 2> gpurun_out/${TAG}_bench_B29600.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fused|k_prep|k_chain|k_update' -s 15 -c 5 -f -o gpurun_out/${TAG}_full python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_fused.log; grep "fused timing" gpurun_out/${TAG}_timing.err | tail -2
python - <<P
import json
for f in ('${TAG}_bench','${TAG}_bench_B29600'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f,'value %.1fM e2e %.1fM ms %.4f frac %.3f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['roofline']['frac']))
    print({k:round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})
P
