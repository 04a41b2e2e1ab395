#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu 2>&1 | tail -8
echo "== gemm table"; timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}.json > $OUT/gemm_${TAG}.txt 2>&1; tail -5 $OUT/gemm_${TAG}.txt
echo "== gemm table sub1"; SCV_TC_SUB=1 timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_sub1.json > $OUT/gemm_${TAG}_sub1.txt 2>&1; tail -5 $OUT/gemm_${TAG}_sub1.txt
echo "== gemm table wsub2"; SCV_TC_WSUB=2 timeout 600 python tools/gemm_bench.py --filter wgrad --json $OUT/gemm_${TAG}_wsub2.json > $OUT/gemm_${TAG}_wsub2.txt 2>&1; tail -3 $OUT/gemm_${TAG}_wsub2.txt
echo "== gemm table pair"; SCV_TC_PAIR=1 timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_pair.json > $OUT/gemm_${TAG}_pair.txt 2>&1; tail -5 $OUT/gemm_${TAG}_pair.txt
echo "== gemm table bf16"; SCV_BENCH_PREC=bf16 timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_bf16.json > $OUT/gemm_${TAG}_bf16.txt 2>&1; tail -5 $OUT/gemm_${TAG}_bf16.txt
for L in enc.3.r3:wgrad enc.2.r3:fwd; do echo "== trace $L"; timeout 300 python tools/tc_trace.py --filter $L 2>&1 | tail -82 > $OUT/trace_${L//[:.]/_}_$TAG.txt; head -34 $OUT/trace_${L//[:.]/_}_$TAG.txt; done
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2', {k:d[k] for k in ('value','ms_per_step')}, d['e2e'].get('value'), d['roofline']['achieved'], d['roofline']['frac'])
for r in d.get('hbm_kernels',[])[:6]: print(r)
"; }
echo "== step tests"; SCV_QUIET=1 timeout 900 python -m pytest tests/test_step_gpu.py -q -m gpu 2>&1 | tail -4
echo "== bench resident"; timeout 900 python bench.py --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -3 $OUT/bench_$TAG.err; show $OUT/bench_$TAG.json resident
echo "== bench c5 resident"; timeout 900 python bench.py --config 5 --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_c5_$TAG.json 2> $OUT/bench_c5_$TAG.err; tail -3 $OUT/bench_c5_$TAG.err; show $OUT/bench_c5_$TAG.json c5
echo "== bench c3"; timeout 900 python bench.py --config 3 --no-gpu-eager --no-cpu --no-sustained > $OUT/bench_c3_$TAG.json 2> $OUT/bench_c3_$TAG.err; tail -3 $OUT/bench_c3_$TAG.err; show $OUT/bench_c3_$TAG.json c3
