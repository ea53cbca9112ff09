#!/bin/bash
# Online serving benchmark on one GPU box: an OpenAI-compatible server (random-init Qwen3-8B) + the
# benchmark_serving client with ShareGPT-shaped synthetic prompts, Poisson arrivals.
#   run_serving_gpu.sh [tp] [rate] [num_prompts]          IMPL=ours (default) | reference ; STAGES=n for staged arrivals
# IMPL=reference starts the UNMODIFIED reference's own server (baseline/_ref, `python -m gllm.entrypoints.api_server`)
# on the same model shape, flags and client: the p50 TTFT / TPOT pair of BASELINE.json's metric for both arms.
TP=${1:-1}; RATE=${2:-16}; N=${3:-300}; PORT=${PORT:-18000}; IMPL=${IMPL:-ours}; STAGES=${STAGES:-1}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${IMPL}_tp${TP}_rate${RATE}
if [ "$IMPL" = "reference" ]; then
  vdir=$(python -c "import importlib.util,os;print(os.path.dirname(importlib.util.find_spec('vllm').origin))")
  export PYTHONPATH=$PWD/baseline/_ref:$PWD/baseline/shims:$PYTHONPATH GLLM_REF_ALIAS_VLLM=1
  [ -f $vdir/_C_stable_libtorch.abi3.so ] && export GLLM_REF_PRELOAD_LIBS=$vdir/_C_stable_libtorch.abi3.so
  MODEL=$(python -c "import sys; sys.path.insert(0,'baseline'); import run_reference, tempfile; print(run_reference.qwen3_8b_dir(tempfile.mkdtemp(prefix='gllm_ref_srv_')))" 2>/dev/null | tail -1)
  setsid python -m gllm.entrypoints.api_server --model-path "$MODEL" --load-format dummy --port $PORT --host 127.0.0.1 \
    --master-addr 127.0.0.1 --master-port $((PORT+1)) --zmq-port-base $((PORT+2)) --tp $TP --maxp 4096 --maxd 1024 \
    --max-cuda-graph-bs 512 --model-max-length 2064 --enable-prefix-caching > gpurun_out/server_$tag.log 2>&1 &
  FMT=words
else
  GLLM_B200_LOG=WARNING setsid python -m gllm_b200.entrypoints.api_server --model-path preset:qwen3-8b --load-format dummy \
    --port $PORT --host 127.0.0.1 --tp $TP --maxp 4096 --maxd 1024 --max-cuda-graph-bs 512 --model-max-length 2064 \
    --enable-prefix-caching > gpurun_out/server_$tag.log 2>&1 &
  FMT=ids
fi
SRV=$!
for i in $(seq 1 360); do
  if curl -s -o /dev/null http://127.0.0.1:$PORT/v1/models; then break; fi
  if ! kill -0 $SRV 2>/dev/null; then echo "server died"; tail -30 gpurun_out/server_$tag.log; exit 1; fi
  sleep 1
done
python benchmarks/benchmark_serving.py --port $PORT --num-prompts $N --request-rate $RATE --max-output-len 512 \
  --prompt-format $FMT --arrival-stage $STAGES --save-result gpurun_out/serving_$tag.json 2>&1 | tr '\r' '\n' | grep -v "it/s" | tail -28
# the server leads its own session (setsid): take down exactly that process group (server + its spawned workers)
kill -TERM -- -$SRV 2>/dev/null
sleep 3
kill -KILL -- -$SRV 2>/dev/null
wait $SRV 2>/dev/null
