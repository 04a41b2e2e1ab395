#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels bnact"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "bnact" 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])
for r in d.get('hbm_kernels',[])[:3]: print(r)
"; }
for O in 1 0 1 0; do echo "== bench spec=$O"; SCV_BNACT_SPEC=$O timeout 900 python bench.py --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_spec${O}_$TAG.json 2> /dev/null; show $OUT/bench_spec${O}_$TAG.json spec$O; done
