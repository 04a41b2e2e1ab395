#!/usr/bin/env bash
# One GPU-box session: kernel tests, step/parity tests, bench, per-layer GEMM table (multicast on / off).
# usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
export SCV_QUIET=0
echo "== kernels"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -15 | tee $OUT/t_kernels_$TAG.log
echo "== kernels (multicast off)"; SCV_TC_MC=0 SCV_TC_WMC=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "not multitile and not split_k" 2>&1 | tail -5 | tee $OUT/t_kernels_nomc_$TAG.log
echo "== step tests"; timeout 900 python -m pytest tests/test_step_gpu.py tests/test_preprocess.py -q -m gpu 2>&1 | tail -15 | tee $OUT/t_step_$TAG.log
echo "== reference parity"; timeout 900 python -m pytest tests/test_reference_gpu.py -q -m gpu 2>&1 | tail -25 | tee $OUT/t_ref_$TAG.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 1500 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
echo "== gemm table (mc on)"; timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_mc.json > $OUT/gemm_${TAG}_mc.txt 2>&1; tail -5 $OUT/gemm_${TAG}_mc.txt
echo "== gemm table (mc off)"; SCV_TC_MC=0 SCV_TC_WMC=0 timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}_nomc.json > $OUT/gemm_${TAG}_nomc.txt 2>&1; tail -5 $OUT/gemm_${TAG}_nomc.txt
