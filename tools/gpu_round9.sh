#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -6
echo "== step+ref tests"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_step_gpu.py tests/test_reference_gpu.py -q -m gpu 2>&1 | tail -12
for L in enc.2.r3:fwd dec.1.skip:fwd enc.3.r0:fwd enc.fc:fwd; do echo "== trace $L"; timeout 300 python tools/tc_trace.py --filter $L 2>&1 | tail -82 > $OUT/trace_${L//[:.]/_}_$TAG.txt; head -44 $OUT/trace_${L//[:.]/_}_$TAG.txt; done
echo "== bench"; timeout 900 python bench.py --no-gpu-eager --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -3 $OUT/bench_$TAG.err
python -c "
import json
d=json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'].get('value'), d['roofline']['achieved'], d['roofline']['frac'])
for r in d['hbm_kernels']: print(r)
"
echo "== bench nofuse"; SCV_FUSE_BNR=0 timeout 900 python bench.py --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_${TAG}_nofuse.json 2> /dev/null
python -c "
import json
d=json.loads(open('$OUT/bench_${TAG}_nofuse.json').read().strip().splitlines()[-1])
print('nofuse', {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['achieved'])
"
