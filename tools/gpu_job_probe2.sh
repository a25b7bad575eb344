#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/peer_gather_probe.py > gpurun_out/peer_probe.txt 2>&1
grep -E "shard|Error|error" gpurun_out/peer_probe.txt | tail -12
