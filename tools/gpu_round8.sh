#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels pair"; SCV_TC_PAIR=1 SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gemm" 2>&1 | tail -4
echo "== gemm table pair"; SCV_TC_PAIR=1 timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_pair.json > $OUT/gemm_${TAG}_pair.txt 2>&1; tail -5 $OUT/gemm_${TAG}_pair.txt
echo "== gemm table single"; timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}.json > $OUT/gemm_${TAG}.txt 2>&1; tail -5 $OUT/gemm_${TAG}.txt
echo "== launch list"; timeout 600 python bench.py --no-graph --steps 3 --warmup 3 --no-gpu-eager --no-cpu --no-sustained > $OUT/plain_$TAG.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $OUT/launches_$TAG.csv python bench.py --no-graph --steps 3 --warmup 3 --no-gpu-eager --no-cpu --no-sustained > $OUT/ncu_$TAG.log 2>&1; tail -2 $OUT/ncu_$TAG.log; wc -l $OUT/launches_$TAG.csv
