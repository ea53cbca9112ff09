#!/bin/bash
# compute-sanitizer passes over the small-shape kernel tests (run on a GPU box; slow).
#   bash tools/sanitize.sh [memcheck|racecheck|synccheck] [pytest -k expression]
TOOL=${1:-memcheck}; EXPR=${2:-"rmsnorm or rope or attn_decode or sample or gemm_bias or mla_attention_decode"}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
compute-sanitizer --tool "$TOOL" --error-exitcode 86 --log-file gpurun_out/sanitizer_$TOOL.log \
  python -m pytest tests/test_kernels_gpu.py -x -q -k "$EXPR" 2>&1 | tail -5
echo "exit=$? (86 = sanitizer errors); summary:"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_$TOOL.log | tail -3
