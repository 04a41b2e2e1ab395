#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== scrubbers vs reference"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_reference_gpu.py -q -m gpu -k "qda or moving_avg" 2>&1 | grep -E "^E|passed|failed" | cut -c1-300 | head -20
echo "== ncu elementwise"; timeout 300 python tools/ew_once.py > $OUT/plain_ew.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "once/" -k regex:"bnact_fwd_kernel|bnact_bwd_apply_kernel|bnact_bwd_reduce_kernel|optim_packed_kernel|recon_loss_chain_kernel|sumsq_packed_kernel|gather_kernel|pack_input_kernel|kl_kernel" -c 48 -o $OUT/prof_ew_$TAG -f python tools/ew_once.py > $OUT/ncu_ew.log 2>&1; tail -2 $OUT/ncu_ew.log
