#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -6
echo "== gemm table"; timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}.json > $OUT/gemm_${TAG}.txt 2>&1; tail -5 $OUT/gemm_${TAG}.txt
echo "== gemm table wsub2"; SCV_TC_WSUB=2 timeout 600 python tools/gemm_bench.py --filter wgrad --json $OUT/gemm_${TAG}_wsub2.json > $OUT/gemm_${TAG}_wsub2.txt 2>&1; tail -3 $OUT/gemm_${TAG}_wsub2.txt
echo "== gemm table sub1"; SCV_TC_SUB=1 timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_sub1.json > $OUT/gemm_${TAG}_sub1.txt 2>&1; tail -5 $OUT/gemm_${TAG}_sub1.txt
echo "== gemm table sub2"; SCV_TC_SUB=2 timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_sub2.json > $OUT/gemm_${TAG}_sub2.txt 2>&1; tail -5 $OUT/gemm_${TAG}_sub2.txt
for L in enc.0.skip:fwd enc.3.skip:fwd; do echo "== trace $L"; timeout 300 python tools/tc_trace.py --filter $L 2>&1 | tail -62 > $OUT/trace_${L//[:.]/_}_$TAG.txt; head -34 $OUT/trace_${L//[:.]/_}_$TAG.txt; done
echo "== bench"; timeout 900 python bench.py --no-gpu-eager --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_'+sys.argv[1]+'.json').read().strip().splitlines()[-1]) if len(sys.argv)>1 else None
PY
python -c "
import json
d=json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'].get('value'), d['roofline']['achieved'], d['roofline']['frac'], d['sustained'])
for r in d['hbm_kernels']: print(r)
"
