"""Print the schemas of the vLLM ops the reference calls (gllm/_custom_ops.py) as they exist in this image."""
import importlib.util
import os
import torch
vdir = os.path.dirname(importlib.util.find_spec('vllm').origin)
import glob
for lib in sorted(glob.glob(vdir + "/*.so") + glob.glob(vdir + "/vllm_flash_attn/*.so")):
    try:
        torch.ops.load_library(lib)
        print("loaded", os.path.basename(lib))
    except Exception as e:  # noqa: BLE001
        print("FAIL", os.path.basename(lib), str(e)[:200])
want = {"_C": ["rms_norm", "fused_add_rms_norm", "rotary_embedding", "batched_rotary_embedding", "silu_and_mul",
               "merge_attn_states", "cutlass_scaled_mm", "dynamic_per_token_scaled_fp8_quant", "per_token_group_fp8_quant"],
        "_C_cache_ops": ["reshape_and_cache_flash", "concat_and_cache_mla", "gather_and_maybe_dequant_cache"],
        "_moe_C": ["topk_softmax", "moe_align_block_size", "moe_sum", "grouped_topk"],
        "_vllm_fa2_C": ["varlen_fwd", "fwd_kvcache"], "_vllm_fa3_C": ["fwd"],
        "_flashmla_C": ["get_mla_decoding_metadata", "fwd_kvcache_mla"]}
for ns, names in want.items():
    for n in names:
        try:
            op = getattr(getattr(torch.ops, ns), n)
            print(ns, n, "::", op.default._schema)
        except Exception as e:  # noqa: BLE001
            print(ns, n, "MISSING", str(e)[:80])
