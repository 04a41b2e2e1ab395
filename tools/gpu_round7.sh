#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== pair pipe"; timeout 120 tools/cu/pair_pipe.bin 2>&1 | tee $OUT/pair_pipe_$TAG.txt
echo "== kernels"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -4
echo "== step+ref tests"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_step_gpu.py tests/test_reference_gpu.py tests/test_preprocess.py -q -m gpu 2>&1 | tail -12
echo "== bench"; timeout 900 python bench.py --no-gpu-eager --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -3 $OUT/bench_$TAG.err
python -c "
import json
d=json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'].get('value'), d['roofline']['achieved'], d['roofline']['frac'])
for r in d['hbm_kernels']: print(r)
"
