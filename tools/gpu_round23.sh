#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== scrubber kernels"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "qda or moving_avg" 2>&1 | grep -E "^E|passed|failed" | cut -c1-300 | head -20
echo "== scrubbers vs reference"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_reference_gpu.py -q -m gpu -k "qda or moving_avg" 2>&1 | grep -E "^E|passed|failed" | cut -c1-300 | head -20
