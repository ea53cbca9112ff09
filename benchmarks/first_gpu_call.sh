#!/bin/bash
# One short single-GPU call that answers the open questions of NOTES.md in priority order (≈4–5 min of box time):
#   gpurun --timeout 600 -- 'bash benchmarks/first_gpu_call.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests";        timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== sync vs async";    timeout 240 bash benchmarks/ab_async.sh 1 2>&1 | tail -2
echo "== reference arm";    timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-700
echo "== tcgen05 prefill attention"; timeout 420 bash benchmarks/validate_attn_tc.sh 2>&1 | tail -12
echo "== MN-major probe";   timeout 90 python benchmarks/umma_mn_sweep.py 2>&1 | tail -6
echo "== memcheck (small)"; timeout 240 bash tools/sanitize.sh memcheck "rmsnorm or rope_mrope or gemm_bias" 2>&1 | tail -4
