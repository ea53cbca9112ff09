"""ChatGLM3 / GLM-4 family: fused HF tensors (query_key_value, dense_h_to_4h), multi-query groups, partial
GPT-J rotary. HF ships this family as remote code, so the check is structural: a ChatGLM checkpoint and the
same weights written under Llama-style names (with the equivalent rotary spec) must generate identical tokens."""
import torch

from gllm_b200.model_loader import ModelLoader
from gllm_b200.models import registry
from gllm_b200.models.decoder import CausalLM
from gllm_b200.models.weight_utils import CheckpointReader


def _glm_cfg():
    return registry.HFConfig({
        "architectures": ["ChatGLMModel"], "hidden_size": 64, "num_layers": 2, "num_attention_heads": 4,
        "kv_channels": 16, "multi_query_attention": True, "multi_query_group_num": 2, "ffn_hidden_size": 128,
        "padded_vocab_size": 256, "layernorm_epsilon": 1e-5, "add_qkv_bias": True, "seq_length": 256,
        "torch_dtype": "float32", "eos_token_id": 2})


def test_chatglm_names_and_partial_rotary():
    torch.manual_seed(0)
    cfg = _glm_cfg()
    spec = registry.spec_chatglm(cfg)
    assert spec.rot_dim == 8 and spec.rope_neox is False and spec.num_kv_heads == 2 and spec.qkv_bias
    h, d, hq, hkv, inter, v = 64, 16, 4, 2, 128, 256
    glm, llama = {}, {}

    def r(*s, std=0.08):
        return torch.randn(*s) * std
    emb, fn, lm = r(v, h, std=0.5), 1 + r(h), r(v, h)
    glm["transformer.embedding.word_embeddings.weight"] = llama["model.embed_tokens.weight"] = emb
    glm["transformer.encoder.final_layernorm.weight"] = llama["model.norm.weight"] = fn
    glm["transformer.output_layer.weight"] = llama["lm_head.weight"] = lm
    for i in range(2):
        g, l = f"transformer.encoder.layers.{i}.", f"model.layers.{i}."
        q, k, vv = r(hq * d, h), r(hkv * d, h), r(hkv * d, h)
        qb, kb, vb = r(hq * d, std=0.3), r(hkv * d, std=0.3), r(hkv * d, std=0.3)
        gate, up, down, o = r(inter, h), r(inter, h), r(h, inter), r(h, hq * d)
        n1, n2 = 1 + r(h), 1 + r(h)
        glm[g + "self_attention.query_key_value.weight"] = torch.cat([q, k, vv])
        glm[g + "self_attention.query_key_value.bias"] = torch.cat([qb, kb, vb])
        glm[g + "self_attention.dense.weight"] = o
        glm[g + "mlp.dense_h_to_4h.weight"] = torch.cat([gate, up])
        glm[g + "mlp.dense_4h_to_h.weight"] = down
        glm[g + "input_layernorm.weight"], glm[g + "post_attention_layernorm.weight"] = n1, n2
        for n, w in (("q", q), ("k", k), ("v", vv)):
            llama[l + f"self_attn.{n}_proj.weight"] = w
        for n, b in (("q", qb), ("k", kb), ("v", vb)):
            llama[l + f"self_attn.{n}_proj.bias"] = b
        llama[l + "self_attn.o_proj.weight"] = o
        llama[l + "mlp.gate_proj.weight"], llama[l + "mlp.up_proj.weight"], llama[l + "mlp.down_proj.weight"] = gate, up, down
        llama[l + "input_layernorm.weight"], llama[l + "post_attention_layernorm.weight"] = n1, n2

    from gllm_b200 import LLM
    prompts = [[5, 17, 99, 200, 3], list(range(10, 70))]
    out = {}
    for name, sd, c in (("glm", glm, cfg),
                        ("llama", llama, registry.HFConfig({
                            "architectures": ["Qwen2ForCausalLM"], "hidden_size": h, "num_hidden_layers": 2,
                            "num_attention_heads": hq, "num_key_value_heads": hkv, "head_dim": d,
                            "intermediate_size": inter, "vocab_size": v, "rms_norm_eps": 1e-5,
                            "max_position_embeddings": 256, "torch_dtype": "float32", "eos_token_id": 2}))):
        llm = LLM(dict(c), load_format="dummy", maxp=32, maxd=16, num_cpu_pages=64, model_max_length=128,
                  log_stats=False, device="cpu")
        model = llm.worker.runner.model
        if name == "llama":  # same rotary convention as GLM: GPT-J pairs on the first half of each head
            from gllm_b200.layers.rotary import build_rope
            rope = build_rope(d, 256, 10000.0, None, d // 2, False)
            model.rope = rope
            for layer in model.layers:
                layer.attn.rope = rope
        model.load_weights(CheckpointReader.from_state_dict(sd))
        out[name] = [s.token_ids for s in llm.generate(tokens=prompts, output_lens=[8, 8], ignore_eos=True)]
        llm.shutdown()
    assert out["glm"] == out["llama"]
