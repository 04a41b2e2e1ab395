#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== mma rate 5 (MN-major)"; timeout 300 tools/cu/mma_rate5.bin 2>&1 | tee $OUT/mma_rate5_$TAG.txt
echo "== kernels"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu 2>&1 | tail -8
for L in enc.3.r3:wgrad dec.0.skip:wgrad enc.2.r3:wgrad dec.2.skip:wgrad enc.3.r3:fwd; do echo "== trace $L"; timeout 300 python tools/tc_trace.py --filter $L 2>&1 | tail -82 > $OUT/trace_${L//[:.]/_}_$TAG.txt; head -60 $OUT/trace_${L//[:.]/_}_$TAG.txt; done
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', {k:d[k] for k in ('value','ms_per_step')}, d['e2e'].get('value'), d['roofline']['achieved'], d['roofline']['frac'])
for r in d.get('hbm_kernels',[])[:6]: print(r)
"; }
echo "== bench resident"; timeout 900 python bench.py --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -3 $OUT/bench_$TAG.err; show $OUT/bench_$TAG.json resident
echo "== bench c5 resident"; timeout 900 python bench.py --config 5 --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_c5_$TAG.json 2> $OUT/bench_c5_$TAG.err; tail -3 $OUT/bench_c5_$TAG.err; show $OUT/bench_c5_$TAG.json c5
