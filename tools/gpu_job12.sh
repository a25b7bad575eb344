#!/bin/bash
TAG=${1:-job12}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_gpu_plugin.py -q -m gpu -x -k "not stress" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/${TAG}_rc.txt; tail -4 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload wikikg2_rotate > gpurun_out/${TAG}_bench_rotate.json 2> gpurun_out/${TAG}_bench_rotate.err
python - <<P
import json
for f in ('${TAG}_bench','${TAG}_bench_rotate'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'value %.1fM e2e %.1fM ms %.4f frac %.3f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['roofline']['frac']))
        print({k:round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})
        if 'beside' in d:
            b=d['beside']; print(' beside: value %.1fM e2e %.1fM ms %.4f'%(b['value']/1e6,b['e2e']['value']/1e6,b['ms_per_step'])); print({k:round(v*1e3,1) for k,v in b['roofline']['kernel_ms'].items()})
    except Exception as e: print(f,'ERR',e)
P
