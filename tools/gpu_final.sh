#!/usr/bin/env bash
# end-of-round validation + evidence: bash tools/gpu_final.sh TAG
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== full gpu suite"; SCV_QUIET=1 timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench default"; timeout 1200 python bench.py > $OUT/bench_final_$TAG.json 2> $OUT/bench_final_$TAG.err; echo "rc=$?"; tail -2 $OUT/bench_final_$TAG.err
echo "== bench reference arm"; timeout 1200 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "rc=$?"; tail -c 600 $OUT/bench_ref_$TAG.json
for C in 3 5 5w201; do echo "== bench config $C"; timeout 900 python bench.py --config $C --no-cpu --no-gpu-eager > $OUT/bench_c${C}_$TAG.json 2> $OUT/bench_c${C}_$TAG.err; echo "rc=$?"; done
python - <<PY
import json
for n in ("final", "c3", "c5", "c5w201"):
    try:
        d = json.loads(open("$OUT/bench_%s_$TAG.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], "e2e", d["e2e"].get("value"), "roof", d["roofline"]["frac"], "sust", (d.get("sustained") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "eager", (d.get("gpu_eager_baseline") or {}).get("value"))
    except Exception as ex:
        print(n, "failed", ex)
PY
echo "== step gaps"; timeout 300 python tools/step_gaps.py 2>&1 | tail -40 > $OUT/step_gaps_$TAG.txt; head -30 $OUT/step_gaps_$TAG.txt
echo "== gemm table"; timeout 600 python tools/gemm_bench.py --json $OUT/gemm_${TAG}.json > $OUT/gemm_${TAG}.txt 2>&1; tail -5 $OUT/gemm_${TAG}.txt
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv python bench.py --no-graph --no-cpu --no-gpu-eager --no-sustained --steps 2 --warmup 3 > $OUT/ncu_launches_$TAG.log 2>&1; echo "rc=$?"; wc -l $OUT/launches_$TAG.csv
echo "== ncu full"; bash tools/gpu_ncu.sh $TAG enc.2.r3:fwd dec.0.skip:dgrad enc.3.r3:wgrad
echo "== ncu elementwise"; timeout 300 python tools/ew_once.py > $OUT/plain_ew.log 2>&1 && timeout 600 ncu --set full --clock-control none --nvtx --nvtx-include "once/" -k regex:"optim_packed_kernel|recon_loss_chain_kernel|sumsq_packed_kernel|pack_input_kernel" -c 4 -o $OUT/prof_ew1_$TAG -f python tools/ew_once.py > $OUT/ncu_ew1.log 2>&1; tail -1 $OUT/ncu_ew1.log
timeout 600 ncu --set full --clock-control none --nvtx --nvtx-include "once/" -k regex:"bnact_fwd_kernel|bnact_bwd_apply_kernel" --launch-skip 14 -c 8 -o $OUT/prof_ew2_$TAG -f python tools/ew_once.py > $OUT/ncu_ew2.log 2>&1; tail -1 $OUT/ncu_ew2.log
ls -la $OUT/prof_ew*_$TAG.ncu-rep; du -sh $OUT
