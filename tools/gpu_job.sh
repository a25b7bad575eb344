#!/bin/bash
# usage: tools/gpu_job.sh <tag> : staged fused-kernel tests, parity suite, short benches; logs under gpurun_out/
TAG=${1:-job}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_dev.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu > gpurun_out/${TAG}_fused.log 2>&1; echo "fused rc=$?" | tee -a gpurun_out/${TAG}_rc.txt
timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_fused.py > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" | tee -a gpurun_out/${TAG}_rc.txt
KGE_B200_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_unfused.json 2> gpurun_out/${TAG}_bench_unfused.err; echo "bench_unfused rc=$?" | tee -a gpurun_out/${TAG}_rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 1000 > gpurun_out/${TAG}_bench_B1000.json 2> gpurun_out/${TAG}_bench_B1000.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 29600 > gpurun_out/${TAG}_bench_B29600.json 2> gpurun_out/${TAG}_bench_B29600.err
tail -3 gpurun_out/${TAG}_fused.log; tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json | head -c 1500
