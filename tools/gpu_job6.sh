#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2 3; do KGE_B200_SPLIT_TRUNC=$m timeout 200 python tools/tf32_round_probe.py 2>&1 | grep "split mode"; done | tee gpurun_out/r2f_round_probe.txt
