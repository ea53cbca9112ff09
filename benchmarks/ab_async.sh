#!/bin/bash
# A/B of synchronous vs asynchronous (lookahead) scheduling on N GPUs of one box; prints one compact line per arm.
#   bash benchmarks/ab_async.sh [N=1] [extra bench.py flags...]
N=${1:-1}; shift || true
cd "$(dirname "$0")/.."
run() {
  if [ "$N" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29741 \
      bench.py --gpus "$N" --steps 1 --warmup 1 "$@"
  else
    python bench.py --steps 1 --warmup 1 "$@"
  fi
}
report() {
  grep "^{" | python -c "
import json, os, sys
d = json.loads(sys.stdin.read())
print(os.environ.get('ARM', ''), 'async' if d['config'].get('async_schedule') else 'sync ', 'tok/s', d['value'],
      'ms/pass', d['ms_per_step'], 'gpu_busy', d['config']['gpu_busy_fraction'], 'p50_tpot', d['latency']['p50_tpot_ms'])"
}
for arm in "" "--async-schedule"; do
  ARM="zmq" run $arm "$@" 2>&1 | ARM="zmq" report
done
if [ "$N" -gt 1 ]; then   # driver -> peer batch fan-out through the shared-memory ring instead of ZeroMQ
  for arm in "" "--async-schedule"; do
    GLLM_BATCH_TRANSPORT=shm run $arm "$@" 2>&1 | ARM="shm" report
  done
fi
