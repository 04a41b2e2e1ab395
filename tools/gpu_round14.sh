#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels+preprocess"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_preprocess.py -q -m gpu 2>&1 | tail -8
echo "== step+ref tests"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_step_gpu.py tests/test_reference_gpu.py -q -m gpu 2>&1 | tail -12
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', {k:d[k] for k in ('value','ms_per_step')}, d['e2e'].get('value'), d['roofline']['achieved'], d['roofline']['frac'])
for r in d.get('hbm_kernels',[])[:8]: print(r)
"; }
echo "== bench resident"; timeout 900 python bench.py --no-gpu-eager --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -3 $OUT/bench_$TAG.err; show $OUT/bench_$TAG.json resident
echo "== bench flat"; timeout 900 python bench.py --no-gpu-eager --no-cpu --no-sustained --no-resident > $OUT/bench_${TAG}_flat.json 2> /dev/null; show $OUT/bench_${TAG}_flat.json flat
echo "== bench c5 resident"; timeout 900 python bench.py --config 5 --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_c5_$TAG.json 2> $OUT/bench_c5_$TAG.err; tail -3 $OUT/bench_c5_$TAG.err; show $OUT/bench_c5_$TAG.json c5
echo "== bench c3 resident"; timeout 900 python bench.py --config 3 --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_c3_$TAG.json 2> $OUT/bench_c3_$TAG.err; tail -3 $OUT/bench_c3_$TAG.err; show $OUT/bench_c3_$TAG.json c3
