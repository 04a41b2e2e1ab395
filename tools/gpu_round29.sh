#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels+step+ref"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py tests/test_reference_gpu.py -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
echo "== bench"; timeout 900 python bench.py --no-gpu-eager --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; python -c "
import json
d=json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['sustained']['value'])
for r in d['hbm_kernels'][:10]: print(r)
"
