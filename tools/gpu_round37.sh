#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels"; SCV_QUIET=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" 2>&1 | tail -2
echo "== step"; SCV_QUIET=1 timeout 300 python -m pytest tests/test_step_gpu.py -q -m gpu 2>&1 | tail -2
for P in auto 0; do echo "== bench pair=$P"; if [ $P = auto ]; then unset SCV_TC_PAIR; else export SCV_TC_PAIR=$P; fi; timeout 300 python bench.py --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_pair${P}_$TAG.json 2> /dev/null; python -c "
import json
d=json.loads(open('$OUT/bench_pair${P}_$TAG.json').read().strip().splitlines()[-1])
print('$P', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['gemm_seconds_per_step'])
"; done
