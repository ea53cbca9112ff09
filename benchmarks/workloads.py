"""Synthetic workloads shaped like the reference's datasets (no network: BASELINE.json asks for
"synthetic ShareGPT-shaped requests"). Optionally reads a local ShareGPT json with --dataset."""
import json
from typing import List, Optional, Tuple

import numpy as np


def sharegpt_shaped(n: int, vocab: int, seed: int = 0, max_prompt: int = 1024, max_total: int = 2048,
                    max_output: Optional[int] = None) -> Tuple[List[List[int]], List[int]]:
    """Log-normal prompt/output lengths clipped by the reference filter (benchmark_throughput.py:41-62:
    prompt >= 4, output >= 4, prompt <= 1024, prompt + output <= 2048; serving: output <= 512)."""
    rng = np.random.default_rng(seed)
    prompts, outs = [], []
    while len(prompts) < n:
        p = int(rng.lognormal(5.0, 1.0))
        o = int(rng.lognormal(5.2, 0.9))
        if max_output is not None:
            o = min(o, max_output)
        if p < 4 or o < 4 or p > max_prompt or p + o > max_total:
            continue
        prompts.append(rng.integers(10, vocab - 10, size=p).tolist())
        outs.append(o)
    return prompts, outs


def from_sharegpt_file(path: str, tokenizer, n: int, seed: int = 0, max_prompt: int = 1024, max_total: int = 2048):
    """First-turn ShareGPT conversations, tokenised and filtered like the reference."""
    with open(path) as f:
        data = json.load(f)
    rng = np.random.default_rng(seed)
    data = [d for d in data if len(d.get("conversations", [])) >= 2]
    rng.shuffle(data)
    prompts, outs = [], []
    for d in data:
        p = tokenizer.encode(d["conversations"][0]["value"])
        o = len(tokenizer.encode(d["conversations"][1]["value"]))
        if len(p) < 4 or o < 4 or len(p) > max_prompt or len(p) + o > max_total:
            continue
        prompts.append(p)
        outs.append(o)
        if len(prompts) == n:
            break
    return prompts, outs


def multi_round_conversations(num_users: int, rounds: int, vocab: int, seed: int = 0, sys_len: int = 256,
                              turn_len: int = 64, answer_len: int = 64):
    """Per-user multi-round conversations sharing a system prompt (prefix-cache benchmark,
    reference: benchmark_prefix_serving.py:69-137)."""
    rng = np.random.default_rng(seed)
    system = rng.integers(10, vocab - 10, size=sys_len).tolist()
    users = []
    for _ in range(num_users):
        turns = [rng.integers(10, vocab - 10, size=int(rng.integers(turn_len // 2, turn_len * 2))).tolist()
                 for _ in range(rounds)]
        users.append(turns)
    return system, users, answer_len
