#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== mma rate 4"; timeout 300 tools/cu/mma_rate4.bin 2>&1 | tee $OUT/mma_rate4_$TAG.txt
for L in dec.0.skip:fwd enc.3.skip:fwd; do echo "== pair trace $L"; SCV_TC_PAIR=1 SCV_TC_MC=0 timeout 300 python tools/tc_trace.py --filter $L 2>&1 | tail -82 > $OUT/trace_pair_${L//[:.]/_}_$TAG.txt; head -70 $OUT/trace_pair_${L//[:.]/_}_$TAG.txt; done
echo "== single trace enc.3.skip"; SCV_TC_MC=0 timeout 300 python tools/tc_trace.py --filter enc.3.skip:fwd 2>&1 | tail -82 > $OUT/trace_single_enc3skip_$TAG.txt; head -60 $OUT/trace_single_enc3skip_$TAG.txt
echo "== device loader test"; timeout 300 python -m pytest tests/test_reference_gpu.py -q -m gpu -k device_loader 2>&1 | tail -5
