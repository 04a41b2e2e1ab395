#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== mma rate"; timeout 300 tools/cu/mma_rate3.bin 2>&1 | tee $OUT/mma_rate3_$TAG.txt
echo "== reference on GPU"; timeout 600 python tools/ref_gpu_profile.py > $OUT/ref_gpu_profile_$TAG.txt 2>&1; head -60 $OUT/ref_gpu_profile_$TAG.txt
for L in enc.0.skip:fwd dec.fc_in:fwd enc.1.r3:fwd dec.0.skip:fwd; do echo "== trace $L"; SCV_TC_MC=0 timeout 300 python tools/tc_trace.py --filter $L 2>&1 | tail -62 > $OUT/trace_${L//[:.]/_}_$TAG.txt; head -50 $OUT/trace_${L//[:.]/_}_$TAG.txt; done
