#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout ${T:-200} "$@" > gpurun_out/pre2_$name.log 2>&1; echo "rc=$?"; grep -E "^\{\"metric|Error|error:|File \"/.*gllm_b200" gpurun_out/pre2_$name.log | cut -c1-${W:-700} | tail -${L:-12}; }
bash tools/tpcheck_variants.sh 2
PORT=29551 run pp2 bash tools/bench_tp.sh 2 --pp 2 --steps 1 --warmup 1 --num-prompts 300
PORT=29561 run pp2_tt bash tools/bench_tp.sh 2 --config llama3-70b-pp4tp2 --layers 8 --steps 1 --warmup 1 --num-prompts 200
CUDA_LAUNCH_BLOCKING=1 PORT=29581 L=30 run dsv3 bash tools/bench_tp.sh 2 --config deepseek-v3-fp8-ep --layers 5 --steps 1 --warmup 1 --num-prompts 100
