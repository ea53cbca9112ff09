#!/usr/bin/env bash
# 2-GPU pre-flight of everything that only ever ran on CPU/gloo: logits-based TP check, PP over NCCL p2p, the other
# BASELINE configurations at reduced depth. Every step is bounded; logs in gpurun_out/pre2_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout ${T:-240} "$@" > gpurun_out/pre2_$name.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pre2_$name.log | cut -c1-${W:-900}; }
GLLM_TP_CHECK_TIMEOUT=200 run tpcheck python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29741 tests/mp_tp_check.py
PORT=29551 run pp2 bash tools/bench_tp.sh 2 --pp 2 --steps 1 --warmup 1 --num-prompts 300
PORT=29561 run pp2_tt bash tools/bench_tp.sh 2 --config llama3-70b-pp4tp2 --layers 8 --steps 1 --warmup 1 --num-prompts 200
PORT=29571 run mixtral bash tools/bench_tp.sh 2 --config mixtral-8x7b-ep --layers 4 --steps 1 --warmup 1 --num-prompts 300
PORT=29581 run dsv3 bash tools/bench_tp.sh 2 --config deepseek-v3-fp8-ep --layers 5 --steps 1 --warmup 1 --num-prompts 200
