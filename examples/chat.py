"""Interactive console chat (reference: examples/chat.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--disable-thinking", action="store_true")
    args = ap.parse_args()
    from gllm_b200 import LLM
    llm = LLM(args.model_path, tp_size=args.tp, pp_size=args.pp, use_thinking=not args.disable_thinking)
    llm.chat()
