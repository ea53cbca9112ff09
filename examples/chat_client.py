"""Streaming chat client against a running server (reference: examples/chat_client.py)."""
import argparse
import json

import requests

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--port", type=int, default=8000)
    args = ap.parse_args()
    url = f"http://127.0.0.1:{args.port}/v1/chat/completions"
    history = []
    while True:
        try:
            q = input(">>> ")
        except EOFError:
            break
        if q.strip() in ("\\quit", "exit"):
            break
        history.append({"role": "user", "content": q})
        answer = ""
        with requests.post(url, json={"messages": history, "stream": True}, stream=True) as r:
            for line in r.iter_lines():
                if line.startswith(b"data: ") and line != b"data: [DONE]":
                    d = json.loads(line[6:])["choices"][0]["delta"].get("content") or ""
                    answer += d
                    print(d, end="", flush=True)
        print()
        history.append({"role": "assistant", "content": answer})
