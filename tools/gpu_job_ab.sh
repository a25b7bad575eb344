#!/bin/bash
# full GPU suite + A/B of the bulk-reduce row scatter on one GPU
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-beside > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
KGE_B200_NO_BULKRED=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-beside > gpurun_out/${TAG}_bench_nobulk.json 2> gpurun_out/${TAG}_bench_nobulk.err
timeout 300 python bench.py --workload freebase_transe_l2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_x.json 2> gpurun_out/${TAG}_bench_x.err
python - <<P
import json
for n in ('bench','bench_nobulk','bench_x'):
    try:
        txt=open('gpurun_out/${TAG}_%s.json'%n).read(); d=json.loads(txt[txt.index('{'):])
        print(n,'value %.1fM e2e %.1fM ms %.4f frac %.3f'%(d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step'],d['roofline']['frac']))
        print({k:round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})
    except Exception as e: print(n,'ERR',e)
P
