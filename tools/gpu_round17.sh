#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
run() { name=$1; shift; env "$@" timeout 600 python tools/gemm_bench.py --filter wgrad --json $OUT/gemm_${TAG}_$name.json > $OUT/gemm_${TAG}_$name.txt 2>&1; echo "$name: $(tail -1 $OUT/gemm_${TAG}_$name.txt)"; }
run base A=1
run st2 SCV_TC_WSTAGES=2
run st3 SCV_TC_WSTAGES=3
run sp1 SCV_TC_WSPLIT=1
run sp3 SCV_TC_WSPLIT=3
run sp4 SCV_TC_WSPLIT=4
run sub2sp2 SCV_TC_WSUB=2 SCV_TC_WSPLIT=2
echo "== trace sub2"; SCV_TC_WSUB=2 timeout 300 python tools/tc_trace.py --filter enc.3.r3:wgrad 2>&1 | tail -82 > $OUT/trace_wsub2_$TAG.txt; head -24 $OUT/trace_wsub2_$TAG.txt
