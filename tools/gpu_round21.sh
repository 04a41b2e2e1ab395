#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -3
run() { name=$1; shift; env "$@" timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_$name.json > $OUT/gemm_${TAG}_$name.txt 2>&1; echo "$name: $(tail -4 $OUT/gemm_${TAG}_$name.txt | tr '\n' ' ')"; }
run base A=1
run rows16 SCV_TC_WROWS=16
run bnk128 SCV_TC_WBNK=128
run bnk192 SCV_TC_WBNK=192
run sp6 SCV_TC_WSPLIT=6
echo "== dram pass"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file $OUT/dram_$TAG.csv python bench.py --no-graph --no-cpu --no-gpu-eager --no-sustained --steps 2 --warmup 3 > $OUT/ncu_dram_$TAG.log 2>&1; echo "rc=$?"
