#!/usr/bin/env bash
# The one 8-GPU call of the round (8 GPU-minutes per minute: every step is tightly bounded, most important first).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout ${T:-120} "$@" > gpurun_out/r8_$name.log 2>&1; echo "rc=$?"; grep -E "^\{|TP_CHECK|NVLS|stall rank|logits rows|MISMATCH|mismatch|Error|error:|File \"/.*gllm_b200" gpurun_out/r8_$name.log | cut -c1-${W:-300} | tail -${L:-12}; }
PORT=29591 T=90 W=3000 L=3 run qwen3_tp8 bash tools/bench_tp.sh 8 --steps 1 --warmup 1
PORT=29561 T=110 W=3000 L=3 run mixtral_ep8 bash tools/bench_tp.sh 8 --config mixtral-8x7b-ep --steps 1 --warmup 1
PORT=29581 T=150 W=3000 L=3 run dsv3_fp8_ep8 bash tools/bench_tp.sh 8 --config deepseek-v3-fp8-ep --layers 16 --steps 1 --warmup 1 --num-prompts 300
PORT=29571 T=160 W=3000 L=3 run llama70b_pp4tp2 bash tools/bench_tp.sh 8 --config llama3-70b-pp4tp2 --tp-mode nccl --steps 1 --warmup 1
GLLM_TP_NVLS=1 GLLM_TP_CHECK_TIMEOUT=100 T=115 run tpcheck8 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29741 tests/mp_tp_check.py
