#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -8
echo "== gemm table dyn"; timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}.json > $OUT/gemm_${TAG}.txt 2>&1; tail -5 $OUT/gemm_${TAG}.txt
echo "== gemm table static"; SCV_TC_DYN=0 timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_static.json > $OUT/gemm_${TAG}_static.txt 2>&1; tail -5 $OUT/gemm_${TAG}_static.txt
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', {k:d[k] for k in ('value','ms_per_step')}, d['e2e'].get('value'), d['roofline']['achieved'], d['roofline']['frac'])
"; }
echo "== step tests"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_step_gpu.py tests/test_reference_gpu.py -q -m gpu 2>&1 | tail -4
echo "== bench dyn"; timeout 900 python bench.py --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -3 $OUT/bench_$TAG.err; show $OUT/bench_$TAG.json dyn
echo "== bench static"; SCV_TC_DYN=0 timeout 900 python bench.py --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_${TAG}_static.json 2> $OUT/bench_${TAG}_static.err; tail -3 $OUT/bench_${TAG}_static.err; show $OUT/bench_${TAG}_static.json static
echo "== bench c5 dyn"; timeout 900 python bench.py --config 5 --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_c5_$TAG.json 2> $OUT/bench_c5_$TAG.err; tail -3 $OUT/bench_c5_$TAG.err; show $OUT/bench_c5_$TAG.json c5
