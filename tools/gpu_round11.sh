#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -4
for C in 5 5w201 3; do
echo "== bench config $C"; timeout 900 python bench.py --config $C --no-gpu-eager --no-cpu > $OUT/bench_c${C}_$TAG.json 2> $OUT/bench_c${C}_$TAG.err; tail -3 $OUT/bench_c${C}_$TAG.err
python -c "
import json
d=json.loads(open('$OUT/bench_c${C}_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','dtype')}, d['e2e'].get('value'), d['roofline'].get('achieved'), d['roofline'].get('frac'), d['roofline'].get('alg_flops_per_step'))
for r in d['hbm_kernels'][:6]: print(r)
"
done
