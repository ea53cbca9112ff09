"""Streaming request client for the serving benchmarks (reference: benchmarks/backend_request_func.py).
Measures per-request TTFT (first non-empty SSE chunk), inter-token latencies and end-to-end latency."""
import json
import time
from dataclasses import dataclass, field
from typing import List, Optional, Union

import aiohttp

AIOHTTP_TIMEOUT = aiohttp.ClientTimeout(total=6 * 60 * 60)


@dataclass
class RequestFuncInput:
    prompt: Union[str, List[int]]
    api_url: str
    prompt_len: int
    output_len: int
    model: str = "default"
    ignore_eos: bool = True
    extra_body: Optional[dict] = None


@dataclass
class RequestFuncOutput:
    generated_text: str = ""
    success: bool = False
    latency: float = 0.0
    output_tokens: int = 0
    ttft: float = 0.0
    itl: List[float] = field(default_factory=list)
    prompt_len: int = 0
    error: str = ""


async def async_request_openai_completions(inp: RequestFuncInput, pbar=None) -> RequestFuncOutput:
    """`/v1/completions` with `temperature 0, top_k 1, ignore_eos, stream` — the payload the reference's
    gllm backend uses (backend_request_func.py:228-238)."""
    payload = {"model": inp.model, "prompt": inp.prompt, "temperature": 0.0, "top_p": 1.0, "top_k": 1,
               "max_tokens": inp.output_len, "stream": True, "ignore_eos": inp.ignore_eos}
    if inp.extra_body:
        payload.update(inp.extra_body)
    out = RequestFuncOutput(prompt_len=inp.prompt_len)
    st = time.perf_counter()
    last = st
    try:
        async with aiohttp.ClientSession(timeout=AIOHTTP_TIMEOUT, trust_env=False) as session:
            async with session.post(url=inp.api_url, json=payload) as resp:
                if resp.status != 200:
                    out.error = f"HTTP {resp.status}: {await resp.text()}"
                    return out
                buf = b""
                async for raw in resp.content.iter_any():
                    buf += raw
                    while b"\n\n" in buf:
                        ev, buf = buf.split(b"\n\n", 1)
                        ev = ev.strip()
                        if not ev.startswith(b"data: "):
                            continue
                        body = ev[6:]
                        if body == b"[DONE]":
                            continue
                        data = json.loads(body)
                        now = time.perf_counter()
                        choice = data["choices"][0] if data.get("choices") else {}
                        text = choice.get("text") or (choice.get("delta") or {}).get("content") or ""
                        if text:
                            if out.ttft == 0.0:
                                out.ttft = now - st
                            else:
                                out.itl.append(now - last)
                            last = now
                            out.generated_text += text
                        if data.get("usage"):
                            out.output_tokens = data["usage"].get("completion_tokens", out.output_tokens)
                out.latency = time.perf_counter() - st
                out.success = True
                if out.output_tokens == 0:
                    out.output_tokens = inp.output_len
    except Exception as e:  # noqa: BLE001
        out.error = repr(e)
    if pbar is not None:
        pbar.update(1)
    return out


async def async_request_openai_chat_completions(inp: RequestFuncInput, pbar=None) -> RequestFuncOutput:
    inp2 = RequestFuncInput(**{**inp.__dict__})
    payload_prompt = inp.prompt if isinstance(inp.prompt, str) else " ".join(map(str, inp.prompt))
    inp2.extra_body = {**(inp.extra_body or {}), "messages": [{"role": "user", "content": payload_prompt}],
                       "max_completion_tokens": inp.output_len}
    return await async_request_openai_completions(inp2, pbar)


ASYNC_REQUEST_FUNCS = {
    "gllm": async_request_openai_completions,
    "gllm_b200": async_request_openai_completions,
    "openai": async_request_openai_completions,
    "openai-chat": async_request_openai_chat_completions,
    "vllm": async_request_openai_completions,
}
