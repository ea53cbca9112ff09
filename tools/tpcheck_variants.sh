#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=${1:-2}
i=0
for v in "GLLM_TP_CHECK_ASYNC=1 GLLM_TP_CHECK_LOGITS=1" "GLLM_TP_CHECK_ASYNC=0 GLLM_TP_CHECK_LOGITS=1" "GLLM_TP_CHECK_ASYNC=1 GLLM_TP_CHECK_LOGITS=0"; do
  i=$((i+1))
  echo "== variant $i: $v"
  env $v GLLM_TP_CHECK_TIMEOUT=${STALL:-75} timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port $((29740+i)) tests/mp_tp_check.py > gpurun_out/tpv_$i.log 2>&1
  echo "rc=$?"
  grep -E "TP_CHECK|stall rank|logits rows|agreement|differ|MISMATCH|mismatch|File \"/.*gllm_b200/(engine|model_runner)" gpurun_out/tpv_$i.log | cut -c1-220 | tail -24
done
