#!/usr/bin/env bash
# Run the reference arm directly (no bench.py wrapper) with its whole stdout/stderr kept: gpurun_out/ref_debug.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
vdir=$(python -c "import importlib.util,os;print(os.path.dirname(importlib.util.find_spec('vllm').origin))")
export PYTHONPATH=$PWD/baseline/_ref:$PWD/baseline/shims:$PYTHONPATH
export GLLM_REF_ALIAS_VLLM=1
[ -f $vdir/_C_stable_libtorch.abi3.so ] && export GLLM_REF_PRELOAD_LIBS=$vdir/_C_stable_libtorch.abi3.so
export TQDM_MININTERVAL=5
timeout ${REF_TIMEOUT:-900} python baseline/run_reference.py --gpus ${1:-1} --steps ${3:-1} --warmup ${4:-1} --num-prompts ${2:-64} > gpurun_out/ref_debug.log 2>&1
echo "rc=$?" >> gpurun_out/ref_debug.log
tail -5 gpurun_out/ref_debug.log
