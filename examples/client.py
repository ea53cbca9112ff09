"""Minimal completions client against a running server (reference: examples/client.py)."""
import argparse
import json

import requests

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--prompt", default="San Francisco is a")
    ap.add_argument("--max-tokens", type=int, default=64)
    ap.add_argument("--stream", action="store_true")
    args = ap.parse_args()
    url = f"http://127.0.0.1:{args.port}/v1/completions"
    body = {"prompt": args.prompt, "max_tokens": args.max_tokens, "stream": args.stream}
    if not args.stream:
        print(requests.post(url, json=body).json()["choices"][0]["text"])
    else:
        with requests.post(url, json=body, stream=True) as r:
            for line in r.iter_lines():
                if line.startswith(b"data: ") and line != b"data: [DONE]":
                    print(json.loads(line[6:])["choices"][0]["text"], end="", flush=True)
        print()
