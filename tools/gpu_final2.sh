#!/usr/bin/env bash
# end-of-round validation (no ncu --set full captures): bash tools/gpu_final2.sh TAG
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
echo "== full gpu suite"; SCV_QUIET=1 timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | cut -c1-220
echo "== bench default"; timeout 1200 python bench.py > $OUT/bench_final_$TAG.json 2> $OUT/bench_final_$TAG.err; echo "rc=$?"
echo "== bench reference arm"; timeout 1200 python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "rc=$?"
for C in 3 5 5w201; do echo "== bench config $C"; timeout 900 python bench.py --config $C --no-cpu --no-gpu-eager > $OUT/bench_c${C}_$TAG.json 2> $OUT/bench_c${C}_$TAG.err; echo "rc=$?"; done
python - <<PY
import json
for n in ("final", "c3", "c5", "c5w201"):
    try:
        d = json.loads(open("$OUT/bench_%s_$TAG.json" % n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], "e2e", d["e2e"].get("value"), "roof", d["roofline"]["frac"], "sust", (d.get("sustained") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "eager", (d.get("gpu_eager_baseline") or {}).get("value"), "launches", d["gpu_launches"])
    except Exception as ex:
        print(n, "failed", ex)
PY
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv python bench.py --no-graph --no-cpu --no-gpu-eager --no-sustained --steps 2 --warmup 3 > $OUT/ncu_launches_$TAG.log 2>&1; echo "rc=$?"; wc -l $OUT/launches_$TAG.csv
echo "== step gaps"; timeout 300 python tools/step_gaps.py 2>&1 | tail -40 > $OUT/step_gaps_$TAG.txt; head -3 $OUT/step_gaps_$TAG.txt | tail -1
