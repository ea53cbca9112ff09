#!/usr/bin/env bash
# Bring-up of the tcgen05 prefill attention kernel (csrc/attn/prefill_attention_tc.cu) in ONE short GPU call:
#   gpurun --timeout 900 -- bash benchmarks/validate_attn_tc.sh
# 1. MN-major B-operand probe: which (LBO, SBO, K-advance) encoding reproduces torch (expected: the CuTe canonical one)
# 2. numerics of the kernel itself vs the fp32 oracle (both KV tile sizes), under a hard timeout: a wrong barrier
#    protocol must not hang the box
# 3. A/B against the mma.sync kernel
# Everything lands in gpurun_out/attn_tc/.
set -uo pipefail
cd "$(dirname "$0")/.."
out=gpurun_out/attn_tc
mkdir -p "$out"
timeout 120 python benchmarks/umma_mn_sweep.py > "$out/mn_sweep.log" 2>&1
echo "mn_sweep rc=$?" | tee -a "$out/summary.txt"
GLLM_ATTN_TC=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "prefill_attention_tc" \
  > "$out/pytest.log" 2>&1
rc=$?
echo "pytest rc=$rc" | tee -a "$out/summary.txt"
tail -5 "$out/pytest.log"
if [[ $rc -ne 0 ]]; then
  # fall back to one N=64 MMA per slab (independent of LBO) to separate descriptor problems from protocol problems
  GLLM_ATTN_TC=1 GLLM_ATTN_TC_V="1024,64,128,1" timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x \
    -k "prefill_attention_tc and 32-8-128-128" > "$out/pytest_split_n.log" 2>&1
  echo "pytest split_n rc=$?" | tee -a "$out/summary.txt"
  exit 0
fi
KB_ATTN_TC=1 timeout 300 python benchmarks/kernel_bench.py attn > "$out/kernel_bench.log" 2>&1
echo "kernel_bench rc=$?" | tee -a "$out/summary.txt"
grep attn_prefill "$out/kernel_bench.log"
