#!/usr/bin/env bash
# ncu --set full captures of single GEMM launches (source-level): bash tools/gpu_ncu.sh <tag> <layer filter> [more filters]
TAG=$1; shift; OUT=gpurun_out; mkdir -p $OUT
for L in "$@"; do
  N=${L//[:.]/_}
  timeout 300 python tools/gemm_bench.py --once --filter $L > $OUT/plain_$N.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "once/" -o $OUT/prof_${N}_$TAG -f python tools/gemm_bench.py --once --filter $L > $OUT/ncu_$N.log 2>&1
  tail -2 $OUT/ncu_$N.log
done
ls -la $OUT/*.ncu-rep | tail -5
