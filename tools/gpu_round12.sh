#!/usr/bin/env bash
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -3
echo "== 2-rank NCCL test"; SCV_QUIET=1 timeout 600 python -m pytest tests/test_parallel_gpu.py -q -m gpu -x 2>&1 | tail -8
echo "== bench n2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_$TAG.json 2> $OUT/bench_n2_$TAG.err; echo "rc=$?"; tail -3 $OUT/bench_n2_$TAG.err
python -c "
import json
d=json.loads(open('$OUT/bench_n2_$TAG.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'].get('value'), d['sustained'])
"
echo "== bench ref arm n2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-400
