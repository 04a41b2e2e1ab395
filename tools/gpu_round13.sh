#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== 2-rank NCCL test"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_parallel_gpu.py -q -m gpu -x 2>&1 | grep -E "^E|passed|failed" | cut -c1-400 | head -30
