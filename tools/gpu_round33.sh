#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "optimizer or gather" 2>&1 | tail -3
echo "== step"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_step_gpu.py -q -m gpu 2>&1 | tail -2
for C in 2 5; do echo "== bench $C"; timeout 900 python bench.py --config $C --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_c${C}_$TAG.json 2> $OUT/bench_c${C}_$TAG.err; python -c "
import json
d=json.loads(open('$OUT/bench_c${C}_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
for r in d['hbm_kernels']:
    if r['kernel'] in ('optim_step','sumsq_packed'): print(r)
"; done
