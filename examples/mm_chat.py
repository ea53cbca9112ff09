"""Multimodal chat requests (text / image / multi-image / video) through the OpenAI-style API
(reference: examples/mm_chat.py). Images are sent as data URLs or http(s)/file URLs."""
import argparse
import base64

import requests


def image_part(path_or_url: str) -> dict:
    if path_or_url.startswith(("http://", "https://", "data:")):
        url = path_or_url
    else:
        with open(path_or_url, "rb") as f:
            url = "data:image/jpeg;base64," + base64.b64encode(f.read()).decode()
    return {"type": "image_url", "image_url": {"url": url}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--image", action="append", default=[])
    ap.add_argument("--video", default=None)
    ap.add_argument("--prompt", default="Describe the picture.")
    args = ap.parse_args()
    content = [image_part(p) for p in args.image]
    if args.video:
        content.append({"type": "video_url", "video_url": {"url": args.video}})
    content.append({"type": "text", "text": args.prompt})
    r = requests.post(f"http://127.0.0.1:{args.port}/v1/chat/completions",
                      json={"messages": [{"role": "user", "content": content}], "max_completion_tokens": 256})
    print(r.json()["choices"][0]["message"]["content"])
