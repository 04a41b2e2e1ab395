#!/usr/bin/env bash
# dynamic vs static work distribution under NCCL overlap: bash tools/gpu_roundN2.sh TAG N
TAG=${1:-x}; N=${2:-2}; OUT=gpurun_out; mkdir -p $OUT
echo "== kernels (many items)"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "many or multitile" 2>&1 | tail -3
port=29560
for C in 2 5; do for D in 1 0; do
  port=$((port+1))
  echo "== bench config $C n$N dyn=$D"
  SCV_TC_DYN=$D timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --config $C --no-cpu --no-gpu-eager > $OUT/bench_c${C}_n${N}_d${D}_$TAG.json 2> $OUT/bench_c${C}_n${N}_d${D}_$TAG.err; echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" $OUT/bench_c${C}_n${N}_d${D}_$TAG.err | tail -3
  python -c "
import json
d=json.loads(open('$OUT/bench_c${C}_n${N}_d${D}_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e'].get('value'), 'sustained', (d.get('sustained') or {}).get('value'), 'roof', d['roofline']['frac'])
"
done; done
echo "== N=1 config 2 dyn / static"
for D in 1 0; do SCV_TC_DYN=$D timeout 600 python bench.py --no-cpu --no-gpu-eager --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dyn=$D', d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
