#!/bin/bash
# Online serving benchmark on one GPU box: api_server (random-init Qwen3-8B) + benchmark_serving client with
# ShareGPT-shaped synthetic token-id prompts, Poisson arrivals. Usage: run_serving_gpu.sh [tp] [rate] [num_prompts]
TP=${1:-1}; RATE=${2:-16}; N=${3:-300}; PORT=18000
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
GLLM_B200_LOG=WARNING python -m gllm_b200.entrypoints.api_server --model-path preset:qwen3-8b --load-format dummy \
  --port $PORT --host 127.0.0.1 --tp $TP --maxp 4096 --maxd 1024 --model-max-length 2064 \
  > gpurun_out/server_tp$TP.log 2>&1 &
SRV=$!
for i in $(seq 1 240); do
  if curl -s -o /dev/null http://127.0.0.1:$PORT/health; then break; fi
  sleep 1
done
python benchmarks/benchmark_serving.py --port $PORT --num-prompts $N --request-rate $RATE --max-output-len 512 \
  --save-result gpurun_out/serving_tp${TP}_rate${RATE}.json 2>&1 | tail -25
kill $SRV
wait $SRV 2>/dev/null
