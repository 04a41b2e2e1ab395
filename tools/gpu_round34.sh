#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "bnact or optimizer" 2>&1 | tail -2
for i in 1 2; do echo "== bench"; timeout 900 python bench.py --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; python -c "
import json
d=json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'])
for r in d['hbm_kernels']:
    if r['kernel'] in ('bnact_fwd','bnact_bwd_apply','sumsq_packed'): print(r)
"; done
