#!/bin/bash
TAG=${1:-n2}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout ${BENCH_TIMEOUT:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
if [ -z "$SKIP_DIST" ]; then timeout 600 python -m pytest tests/test_dist.py -q -m gpu -x > gpurun_out/${TAG}_dist.log 2>&1; tail -3 gpurun_out/${TAG}_dist.log; fi
python - <<P
import json
try:
    txt=open('gpurun_out/${TAG}_bench.json').read(); d=json.loads(txt[txt.index('{'):])
    print('N=%d value %.1fM e2e %.1fM ms %.4f'%(d['n_gpus'],d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step']), d['config']['workload'][:50])
    print({k:round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})
    if 'beside' in d:
        b=d['beside']; print(' beside: value %.1fM e2e %.1fM ms %.4f'%(b['value']/1e6,b['e2e']['value']/1e6,b['ms_per_step'])); print({k:round(v*1e3,1) for k,v in b['roofline']['kernel_ms'].items()})
except Exception as e: print('ERR',e)
P
