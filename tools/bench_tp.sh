#!/usr/bin/env bash
# bench.py at N GPUs the way the driver launches it: tools/bench_tp.sh N [extra bench.py args...] (env passes through)
n=$1; shift
cd "$(dirname "$0")/.."
if [ "$n" = "1" ]; then exec python bench.py --gpus 1 "$@"; fi
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port ${PORT:-29500} bench.py --gpus "$n" "$@"
