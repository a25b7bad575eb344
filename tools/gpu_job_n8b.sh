#!/bin/bash
# 8 GPUs: the default bench line (Freebase shape + FB15k beside), then the Freebase shape again with NCCL capped at 16 CTAs
TAG=${1:-n8b}
N=${2:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5"
timeout 300 $RUN > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
NCCL_MAX_CTAS=16 timeout 200 $RUN --no-beside > gpurun_out/${TAG}_bench_cta16.json 2> gpurun_out/${TAG}_bench_cta16.err; echo "bench rc=$?"
python - <<P
import json
for n in ('bench','bench_cta16'):
    try:
        txt=open('gpurun_out/${TAG}_%s.json'%n).read(); d=json.loads(txt[txt.index('{'):])
        print(n,'N=%d value %.1fM e2e %.1fM ms %.4f'%(d['n_gpus'],d['value']/1e6,d['e2e']['value']/1e6,d['ms_per_step']), d['config']['workload'][:40])
        print({k:round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})
        if 'beside' in d:
            b=d['beside']; print(' beside: value %.1fM e2e %.1fM ms %.4f'%(b['value']/1e6,b['e2e']['value']/1e6,b['ms_per_step'])); print({k:round(v*1e3,1) for k,v in b['roofline']['kernel_ms'].items()})
    except Exception as e: print(n,'ERR',e)
P
