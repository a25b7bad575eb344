#!/bin/bash
TAG=${1:-d1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py tests/test_gpu_plugin.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
