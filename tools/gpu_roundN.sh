#!/usr/bin/env bash
# multi-GPU round: bash tools/gpu_roundN.sh TAG N [configs...]
TAG=${1:-x}; N=${2:-2}; shift 2; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then echo "== 2-rank NCCL test"; SCV_QUIET=1 timeout 500 python -m pytest tests/test_parallel_gpu.py -q -m gpu -x 2>&1 | grep -E "^E|passed|failed" | cut -c1-300 | head; fi
port=29540
for C in "$@"; do
  port=$((port+1))
  echo "== bench config $C n$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --config $C --no-cpu --no-gpu-eager > $OUT/bench_c${C}_n${N}_$TAG.json 2> $OUT/bench_c${C}_n${N}_$TAG.err; echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" $OUT/bench_c${C}_n${N}_$TAG.err | tail -4
  python -c "
import json
d=json.loads(open('$OUT/bench_c${C}_n${N}_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e'].get('value'), 'sustained', (d.get('sustained') or {}).get('value'), 'roof', d['roofline']['frac'])
"
done
