#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== ref test dyn"; SCV_QUIET=1 timeout 900 python -m pytest "tests/test_reference_gpu.py::test_benchmarked_config_against_reference_on_gpu" -q -m gpu 2>&1 | grep -E "^E|passed|failed|assert" | cut -c1-400 | head -30
echo "== ref test static"; SCV_TC_DYN=0 SCV_QUIET=1 timeout 900 python -m pytest "tests/test_reference_gpu.py::test_benchmarked_config_against_reference_on_gpu" -q -m gpu 2>&1 | grep -E "^E|passed|failed|assert" | cut -c1-400 | head -30
echo "== ref test static wsplit2"; SCV_TC_WSPLIT=2 SCV_TC_DYN=0 SCV_QUIET=1 timeout 900 python -m pytest "tests/test_reference_gpu.py::test_benchmarked_config_against_reference_on_gpu" -q -m gpu 2>&1 | grep -E "^E|passed|failed|assert" | cut -c1-400 | head -30
