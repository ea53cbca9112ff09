"""Build one global random HF-style state dict for a tiny config and load it through the normal
sharding path (`model.load_weights`) — so tp/pp/ep layouts of the same model hold the same weights."""
import torch

from gllm_b200.models.weight_utils import CheckpointReader


def global_state_dict(cfg: dict, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    h, L = cfg["hidden_size"], cfg["num_hidden_layers"]
    hq, hkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    d = cfg.get("head_dim") or h // hq
    inter, v = cfg["intermediate_size"], cfg["vocab_size"]
    arch = cfg["architectures"][0]

    def r(*shape, std=0.05):
        return torch.randn(*shape, generator=g) * std

    sd = {"model.embed_tokens.weight": r(v, h, std=0.5), "model.norm.weight": 1 + r(h, std=0.1),
          "lm_head.weight": r(v, h)}
    for i in range(L):
        p = f"model.layers.{i}."
        sd[p + "input_layernorm.weight"] = 1 + r(h, std=0.1)
        sd[p + "post_attention_layernorm.weight"] = 1 + r(h, std=0.1)
        sd[p + "self_attn.q_proj.weight"] = r(hq * d, h)
        sd[p + "self_attn.k_proj.weight"] = r(hkv * d, h)
        sd[p + "self_attn.v_proj.weight"] = r(hkv * d, h)
        sd[p + "self_attn.o_proj.weight"] = r(h, hq * d)
        if arch in ("Qwen2ForCausalLM",):
            for n, sz in (("q", hq * d), ("k", hkv * d), ("v", hkv * d)):
                sd[p + f"self_attn.{n}_proj.bias"] = r(sz, std=0.2)
        if arch in ("Qwen3ForCausalLM",):
            sd[p + "self_attn.q_norm.weight"] = 1 + r(d, std=0.1)
            sd[p + "self_attn.k_norm.weight"] = 1 + r(d, std=0.1)
        if arch == "MixtralForCausalLM":
            e = cfg["num_local_experts"]
            sd[p + "block_sparse_moe.gate.weight"] = r(e, h, std=0.5)
            for j in range(e):
                q = p + f"block_sparse_moe.experts.{j}."
                sd[q + "w1.weight"], sd[q + "w3.weight"], sd[q + "w2.weight"] = r(inter, h), r(inter, h), r(h, inter)
        else:
            sd[p + "mlp.gate_proj.weight"] = r(inter, h)
            sd[p + "mlp.up_proj.weight"] = r(inter, h)
            sd[p + "mlp.down_proj.weight"] = r(h, inter)
    return sd


def load_global_weights(model, cfg: dict, seed: int = 0):
    model.load_weights(CheckpointReader.from_state_dict(global_state_dict(cfg, seed)))
