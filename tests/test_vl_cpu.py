"""Qwen2.5-VL / Qwen3-VL / Qwen3-VL-MoE against HuggingFace on tiny random models (CPU, fp32): vision tower,
placeholder merge, M-RoPE positions (chunked + interleaved), DeepStack, chunked prefill across an image."""
from conftest import scratch_dir

import numpy as np
import pytest
import torch

transformers = pytest.importorskip("transformers")

IMG, VID, VSTART = 290, 291, 292


def _pixels(grids, patch, tps=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    n = sum(t * h * w for t, h, w in grids)
    return torch.randn(n, 3 * tps * patch * patch, generator=g)


def _prompt(grids, merge=2):
    ids = [5, 17, 99]
    for t, h, w in grids:
        ids += [VSTART] + [IMG] * (t * (h // merge) * (w // merge)) + [7, 8]
    return ids + [45, 46, 47]


def _hf_greedy(model, ids, pix, grids, n):
    types = torch.tensor([[1 if t == IMG else 0 for t in ids]])   # HF needs it to build the M-RoPE index
    with torch.no_grad():
        out = model.generate(input_ids=torch.tensor([ids]), pixel_values=pix, image_grid_thw=torch.tensor(grids),
                             mm_token_type_ids=types, max_new_tokens=n, do_sample=False, eos_token_id=None, pad_token_id=0)
    return out[0, len(ids):].tolist()


def _ours(model, ids, pix, grids, n, **kw):
    from gllm_b200 import LLM
    d = scratch_dir("gllm_b200_vl_")
    model.save_pretrained(d, safe_serialization=True)
    args = dict(maxp=64, maxd=64, page_size=16, num_cpu_pages=96, model_max_length=320, log_stats=False)
    args.update(kw)
    llm = LLM(d, **args)
    mm = {"pixel_values": pix.numpy(), "image_grid_thw": np.asarray(grids)}
    text_only = [5, 17, 99, 200, 3]
    outs = llm.generate(tokens=[ids, text_only], output_lens=[n, n], ignore_eos=True, mm_contents=[mm, None])
    llm.shutdown()
    return outs[0].token_ids[len(ids):], outs[1].token_ids[len(text_only):], text_only


def test_qwen2_5_vl_matches_hf():
    from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration
    torch.manual_seed(0)
    cfg = Qwen2_5_VLConfig(
        text_config=dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                         intermediate_size=128, vocab_size=300, max_position_embeddings=512,
                         rope_parameters={"rope_type": "default", "mrope_section": [2, 3, 3], "rope_theta": 10000.0}),
        vision_config=dict(depth=2, hidden_size=32, num_heads=2, intermediate_size=64, out_hidden_size=64,
                           patch_size=14, spatial_merge_size=2, temporal_patch_size=2, window_size=56,
                           fullatt_block_indexes=[1], in_channels=3),
        image_token_id=IMG, video_token_id=VID, vision_start_token_id=VSTART, tie_word_embeddings=False)
    model = Qwen2_5_VLForConditionalGeneration(cfg).eval().float()
    grids = [(1, 6, 4), (1, 4, 8)]          # two images; 6x4 patches -> windows of 2x2 merged tokens
    pix = _pixels(grids, 14)
    ids = _prompt(grids)
    ref = _hf_greedy(model, ids, pix, grids, 8)
    got, got_text, text_only = _ours(model, ids, pix, grids, 8)
    assert got == ref
    with torch.no_grad():
        ref_text = model.generate(torch.tensor([text_only]), max_new_tokens=8, do_sample=False, eos_token_id=None,
                                  pad_token_id=0)[0, len(text_only):].tolist()
    assert got_text == ref_text
    # chunked prefill cutting through the first image (maxp=8 tokens per step)
    got2, _, _ = _ours(model, ids, pix, grids, 8, maxp=8)
    assert got2 == ref


def _qwen3_vl_cfg(moe: bool):
    text = dict(hidden_size=64, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
                intermediate_size=128, vocab_size=300, max_position_embeddings=512,
                rope_parameters={"rope_type": "default", "mrope_section": [2, 3, 3], "mrope_interleaved": True,
                                 "rope_theta": 10000.0})
    if moe:
        text.update(num_experts=4, num_experts_per_tok=2, moe_intermediate_size=32, decoder_sparse_step=1,
                    mlp_only_layers=[], norm_topk_prob=True)
    vision = dict(depth=3, hidden_size=32, num_heads=2, intermediate_size=64, out_hidden_size=64, patch_size=16,
                  spatial_merge_size=2, temporal_patch_size=2, num_position_embeddings=16,
                  deepstack_visual_indexes=[0, 1], in_channels=3)
    return dict(text_config=text, vision_config=vision, image_token_id=IMG, video_token_id=VID,
                vision_start_token_id=VSTART, tie_word_embeddings=False)


def test_qwen3_vl_deepstack_matches_hf():
    from transformers import Qwen3VLConfig, Qwen3VLForConditionalGeneration
    torch.manual_seed(1)
    model = Qwen3VLForConditionalGeneration(Qwen3VLConfig(**_qwen3_vl_cfg(False))).eval().float()
    grids = [(1, 4, 6)]
    pix = _pixels(grids, 16)
    ids = _prompt(grids)
    ref = _hf_greedy(model, ids, pix, grids, 8)
    got, _, _ = _ours(model, ids, pix, grids, 8)
    assert got == ref
    got2, _, _ = _ours(model, ids, pix, grids, 8, maxp=8)   # DeepStack rows split across prefill chunks
    assert got2 == ref


def test_qwen3_vl_moe_matches_hf():
    from transformers import Qwen3VLMoeConfig, Qwen3VLMoeForConditionalGeneration
    torch.manual_seed(2)
    model = Qwen3VLMoeForConditionalGeneration(Qwen3VLMoeConfig(**_qwen3_vl_cfg(True))).eval().float()
    grids = [(1, 4, 4)]
    pix = _pixels(grids, 16)
    ids = _prompt(grids)
    ref = _hf_greedy(model, ids, pix, grids, 6)
    got, _, _ = _ours(model, ids, pix, grids, 6)
    assert got == ref


def test_mrope_positions_text_image_text():
    from gllm_b200.models.multimodal import MMInfo, compute_mrope_positions
    info = MMInfo(image_token_id=IMG, video_token_id=VID, spatial_merge_size=2)
    ids = [1, 2, VSTART] + [IMG] * 6 + [3, 4]      # grid 1 x 4 x 6 -> 2 x 3 merged
    pos, delta = compute_mrope_positions(ids, [(1, 4, 6)], None, info)
    assert pos[:, :3].tolist() == [[0, 1, 2]] * 3
    assert pos[0, 3:9].tolist() == [3] * 6
    assert pos[1, 3:9].tolist() == [3, 3, 3, 4, 4, 4]
    assert pos[2, 3:9].tolist() == [3, 4, 5, 3, 4, 5]
    assert pos[:, 9:].tolist() == [[6, 7]] * 3       # continues at max + 1
    assert delta == 8 - len(ids)


def test_mrope_positions_video_temporal_spacing():
    """Qwen2.5-VL video: one placeholder run of t*h*w tokens, temporal index = frame * second_per_grid *
    tokens_per_second (reference: gllm/layers/rotary_embedding.py:697-707); Qwen3-VL: one run per frame."""
    from gllm_b200.models.multimodal import MMInfo, compute_mrope_positions
    info = MMInfo(image_token_id=IMG, video_token_id=VID, spatial_merge_size=2, tokens_per_second=2.0)
    ids = [1, VSTART] + [VID] * 12 + [3]          # grid 3 x 4 x 4 -> 3 frames of 2 x 2 merged tokens
    pos, delta = compute_mrope_positions(ids, None, [(3, 4, 4)], info, second_per_grid_ts=[1.0])
    assert pos[0, 2:14].tolist() == [2] * 4 + [4] * 4 + [6] * 4           # start 2, frames 0/2/4 apart
    assert pos[1, 2:14].tolist() == [2, 2, 3, 3] * 3 and pos[2, 2:14].tolist() == [2, 3, 2, 3] * 3
    assert pos[:, 14].tolist() == [7, 7, 7] and delta == 8 - len(ids)
    info3 = MMInfo(image_token_id=IMG, video_token_id=VID, spatial_merge_size=2, per_frame_video=True)
    ids3 = [1] + [VID] * 4 + [9] + [VID] * 4 + [3]   # two frames of one video, a timestamp token in between
    pos3, _ = compute_mrope_positions(ids3, None, [(2, 4, 4)], info3)
    assert pos3[0, 1:5].tolist() == [1] * 4 and pos3[1, 1:5].tolist() == [1, 1, 2, 2]
    assert pos3[:, 5].tolist() == [3, 3, 3] and pos3[0, 6:10].tolist() == [4] * 4
