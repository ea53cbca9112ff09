"""Accuracy sanity check: 5-shot CoT MMLU-Pro through /v1/chat/completions
(reference: benchmarks/evaluate_MMLU_pro.py). Needs a LOCAL copy of the dataset
(`--data-dir` with test/validation parquet or json files — there is no network here).

    python benchmarks/evaluate_MMLU_pro.py --data-dir /data/MMLU-Pro --port 8000 --num-per-subject 100
"""
import argparse
import glob
import json
import os
import re
from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor

import requests

CHOICES = "ABCDEFGHIJ"


def load_split(data_dir, split):
    files = sorted(glob.glob(os.path.join(data_dir, "**", f"{split}*.parquet"), recursive=True))
    if files:
        import pandas as pd
        return pd.concat([pd.read_parquet(f) for f in files]).to_dict("records")
    files = sorted(glob.glob(os.path.join(data_dir, "**", f"{split}*.json*"), recursive=True))
    rows = []
    for f in files:
        with open(f) as fh:
            rows += [json.loads(l) for l in fh] if f.endswith("l") else json.load(fh)
    if not rows:
        raise FileNotFoundError(f"no {split} split under {data_dir}")
    return rows


def fmt(q, with_answer):
    s = "Question: " + q["question"] + "\nOptions:\n"
    for i, o in enumerate(q["options"]):
        s += f"{CHOICES[i]}. {o}\n"
    if with_answer:
        cot = q.get("cot_content", "").replace("A: Let's think step by step.", "Answer: Let's think step by step.")
        s += (cot or f"Answer: The answer is ({q['answer']}).") + "\n\n"
    else:
        s += "Answer: Let's think step by step."
    return s


def extract(text):
    m = re.search(r"answer is \(?([A-J])\)?", text)
    if m:
        return m.group(1)
    m = re.findall(r"\b([A-J])\b", text)
    return m[-1] if m else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--num-per-subject", type=int, default=100)
    ap.add_argument("--max-tokens", type=int, default=1024)
    ap.add_argument("--workers", type=int, default=64)
    args = ap.parse_args()
    test, val = load_split(args.data_dir, "test"), load_split(args.data_dir, "validation")
    shots = defaultdict(list)
    for q in val:
        shots[q["category"]].append(q)
    per = defaultdict(list)
    for q in test:
        if len(per[q["category"]]) < args.num_per_subject:
            per[q["category"]].append(q)
    url = f"http://{args.host}:{args.port}/v1/chat/completions"

    def ask(q):
        prompt = ("The following are multiple choice questions (with answers) about "
                  f"{q['category']}. Think step by step and then finish your answer with "
                  "\"the answer is (X)\" where X is the correct letter choice.\n\n")
        prompt += "".join(fmt(s, True) for s in shots[q["category"]][:5]) + fmt(q, False)
        r = requests.post(url, json={"messages": [{"role": "user", "content": prompt}],
                                     "max_completion_tokens": args.max_tokens, "temperature": 0.0, "top_k": 1})
        return extract(r.json()["choices"][0]["message"]["content"]) == q["answer"]

    total = correct = 0
    for cat, qs in sorted(per.items()):
        with ThreadPoolExecutor(args.workers) as ex:
            res = list(ex.map(ask, qs))
        total += len(res)
        correct += sum(res)
        print(f"{cat:<20} {sum(res) / len(res):.3f} ({len(res)})")
    print(f"{'overall':<20} {correct / max(total, 1):.3f} ({total})")


if __name__ == "__main__":
    main()
