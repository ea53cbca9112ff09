"""OpenAI-compatible HTTP server (reference: gllm/entrypoints/api_server.py:36-322).

    python -m gllm_b200.entrypoints.api_server --model-path <hf dir | preset:name> [--tp 8 --pp 1 ...]

Routes kept from the reference: GET /v1/models, POST /v1/chat/completions, POST /v1/completions,
POST /start_profile, POST /stop_profile. Added: GET /health, GET /metrics (Prometheus text),
POST /tokenize, POST /detokenize; `usage` is filled in (the reference always returns zeros).
Streaming = server-sent events `data: {json}\\n\\n` ... `data: [DONE]\\n\\n`; a client disconnect
aborts the request and frees its KV pages.
"""
from __future__ import annotations

import argparse
import asyncio
from http import HTTPStatus
from typing import List, Optional

from fastapi import Request  # module level: FastAPI resolves the (string) annotations of the handlers here

from gllm_b200.entrypoints.protocol import (ChatCompletionRequest, ChatCompletionResponse,
                                            ChatCompletionResponseChoice, ChatCompletionResponseStreamChoice,
                                            ChatCompletionStreamResponse, ChatMessage, CompletionRequest,
                                            CompletionResponse, CompletionResponseChoice,
                                            CompletionResponseStreamChoice, CompletionStreamResponse, DeltaMessage,
                                            DetokenizeRequest, DetokenizeResponse, ErrorResponse, ModelCard, ModelList,
                                            ModelPermission, TokenizeRequest, TokenizeResponse, UsageInfo)
from gllm_b200.utils.logging import logger

llm = None  # AsyncLLM, set by build_app / main


def _error(msg: str, status=HTTPStatus.BAD_REQUEST):
    from fastapi.responses import JSONResponse
    return JSONResponse(ErrorResponse(message=msg, type="BadRequestError", code=status.value).model_dump(),
                        status_code=status.value)


def _validate_sampling(request):
    if request.temperature is not None and request.temperature < 0:
        return "temperature must be >= 0"
    if request.top_p is not None and not 0.0 < request.top_p <= 1.0:
        return "top_p must be in (0, 1]"
    if request.repetition_penalty is not None and request.repetition_penalty <= 0:
        return "repetition_penalty must be > 0"
    return None


def _validate(token_ids, output_len, vocab_size=None):
    """-> error message or None. An empty prompt or a token id outside the vocabulary would take the whole engine
    down (there is no row to sample from / the embedding lookup faults), so they are refused at the door."""
    if not token_ids:
        return "the prompt is empty"
    if output_len is not None and output_len < 1:
        return "max_tokens must be at least 1"
    if vocab_size is not None and (min(token_ids) < 0 or max(token_ids) >= vocab_size):
        return f"token ids must be in [0, {vocab_size})"
    if not llm.check_seq_length(token_ids, output_len):
        return "seq length exceeds max model length"
    return None


def _usage(stream) -> UsageInfo:
    return UsageInfo(prompt_tokens=stream.prompt_tokens, completion_tokens=stream.completion_tokens,
                     total_tokens=stream.prompt_tokens + stream.completion_tokens)


# ------------------------------------------------------------------------------------------------
# response generators (reference: serving_chat.py / serving_completions.py)
# ------------------------------------------------------------------------------------------------
async def chat_completion_generator(stream, request) -> ChatCompletionResponse:
    text = await llm.collect(stream)
    choice = ChatCompletionResponseChoice(index=0, message=ChatMessage(role="assistant", content=text),
                                          finish_reason=stream.finish_reason)
    return ChatCompletionResponse(choices=[choice], usage=_usage(stream), model=request.model)


async def chat_completion_stream_generator(stream, request):
    rid = None
    first = True
    try:
        async for delta in stream:
            dm = DeltaMessage(role="assistant", content=delta) if first else DeltaMessage(content=delta)
            first = False
            chunk = ChatCompletionStreamResponse(
                choices=[ChatCompletionResponseStreamChoice(index=0, delta=dm)], model=request.model)
            if rid is None:
                rid = chunk.id
            chunk.id = rid
            yield f"data: {chunk.model_dump_json(exclude_none=True)}\n\n"
    finally:
        llm.abort_stream(stream)  # no-op unless the client disconnected mid-stream
    final = ChatCompletionStreamResponse(
        choices=[ChatCompletionResponseStreamChoice(index=0, delta=DeltaMessage(),
                                                    finish_reason=stream.finish_reason or "stop")],
        model=request.model, usage=_usage(stream))
    if rid is not None:
        final.id = rid
    yield f"data: {final.model_dump_json(exclude_none=True)}\n\n"
    yield "data: [DONE]\n\n"


async def completion_generator(stream, request) -> CompletionResponse:
    text = await llm.collect(stream)
    choice = CompletionResponseChoice(index=0, text=text, finish_reason=stream.finish_reason)
    return CompletionResponse(choices=[choice], model=request.model, usage=_usage(stream))


async def completion_stream_generator(stream, request):
    rid = None
    try:
        async for delta in stream:
            chunk = CompletionStreamResponse(choices=[CompletionResponseStreamChoice(index=0, text=delta)],
                                             model=request.model)
            if rid is None:
                rid = chunk.id
            chunk.id = rid
            yield f"data: {chunk.model_dump_json(exclude_unset=False)}\n\n"
    finally:
        llm.abort_stream(stream)  # no-op unless the client disconnected mid-stream
    final = CompletionStreamResponse(
        choices=[CompletionResponseStreamChoice(index=0, text="", finish_reason=stream.finish_reason or "stop")],
        model=request.model, usage=_usage(stream))
    if rid is not None:
        final.id = rid
    yield f"data: {final.model_dump_json(exclude_unset=False)}\n\n"
    yield "data: [DONE]\n\n"


# ------------------------------------------------------------------------------------------------
def _encode_prompt(prompt) -> List[int]:
    if isinstance(prompt, str):
        return llm.encode(prompt)
    if isinstance(prompt, list) and prompt and isinstance(prompt[0], int):
        return list(prompt)
    if isinstance(prompt, list) and prompt and isinstance(prompt[0], list):
        return list(prompt[0])
    if isinstance(prompt, list) and prompt and isinstance(prompt[0], str):
        return llm.encode(prompt[0])
    raise ValueError("unsupported prompt type")


def build_app(engine):
    """FastAPI app bound to an AsyncLLM (also used by the tests with fastapi.testclient)."""
    import fastapi
    from fastapi.responses import JSONResponse, PlainTextResponse, StreamingResponse
    global llm
    llm = engine
    app = fastapi.FastAPI(title="gllm_b200")
    async def _in_thread(fn, *a, **kw):
        return await asyncio.get_running_loop().run_in_executor(None, lambda: fn(*a, **kw))

    @app.get("/health")
    async def health():
        llm.check_worker_alive()
        if llm.failed:
            return JSONResponse({"status": "engine failure", "detail": llm.failed}, status_code=500)
        return JSONResponse({"status": "ok"})

    @app.get("/metrics")
    async def metrics():
        m, s = llm.metrics, llm.last_stats or {}
        lines = [
            "# TYPE gllm_requests_total counter", f"gllm_requests_total {m['requests_total']}",
            "# TYPE gllm_requests_finished_total counter", f"gllm_requests_finished_total {m['requests_finished']}",
            "# TYPE gllm_requests_aborted_total counter", f"gllm_requests_aborted_total {m['requests_aborted']}",
            "# TYPE gllm_prompt_tokens_total counter", f"gllm_prompt_tokens_total {m['prompt_tokens_total']}",
            "# TYPE gllm_generation_tokens_total counter", f"gllm_generation_tokens_total {m['generation_tokens_total']}",
            *llm.hist["ttft"].lines("gllm_time_to_first_token_seconds"),
            *llm.hist["tpot"].lines("gllm_time_per_output_token_seconds"),
            *llm.hist["e2e"].lines("gllm_e2e_request_latency_seconds"),
            "# TYPE gllm_num_requests_running gauge", f"gllm_num_requests_running {len(llm.running_maps)}",
            "# TYPE gllm_num_requests_waiting gauge", f"gllm_num_requests_waiting {s.get('wait', 0)}",
            "# TYPE gllm_kv_cache_usage_perc gauge", f"gllm_kv_cache_usage_perc {s.get('memory_util', 0.0)}",
            "# TYPE gllm_prefix_cache_hit_rate gauge", f"gllm_prefix_cache_hit_rate {s.get('cache_hit_rate', 0.0)}",
            "# TYPE gllm_num_preemptions_total counter", f"gllm_num_preemptions_total {s.get('preempted', 0)}",
        ]
        # per-phase step accounting from the driver worker (launch -> tokens-ready time of every micro-batch;
        # a batch that carries any prefill chunk counts as prefill)
        for key, kind in (("seconds", "seconds"), ("count", "iterations"), ("tokens", "tokens")):
            lines.append(f"# TYPE gllm_step_{kind}_total counter")
            for ph in ("prefill", "decode"):
                lines.append(f'gllm_step_{kind}_total{{phase="{ph}"}} {s.get(f"{ph}_step_{key}", 0):.6g}')
        return PlainTextResponse("\n".join(lines) + "\n")

    @app.get("/v1/models")
    async def show_available_models():
        name = str(llm.cfg.model_path)
        models = ModelList(data=[ModelCard(id=name, root=name, max_model_len=llm.model_max_length,
                                           permission=[ModelPermission()])])
        return JSONResponse(content=models.model_dump())

    @app.post("/v1/chat/completions")
    async def create_chat_completion(request: ChatCompletionRequest, raw_request: Request):
        mm_contents = None
        try:
            if llm.loader.use_mm:
                from gllm_b200.models.multimodal import encode_mm
                token_ids, mm_contents = await _in_thread(encode_mm, llm, request.messages)
            else:
                token_ids = await _in_thread(llm.encode, None, True, request.messages)
        except Exception as e:  # noqa: BLE001
            return _error(f"cannot encode messages: {e}")
        bad = _validate_sampling(request) or _validate(token_ids, request.output_len(),
                                                       llm.loader.config.get("vocab_size"))
        if bad:
            return _error(bad)
        if llm.failed:
            return _error(f"engine is down: {llm.failed}", HTTPStatus.INTERNAL_SERVER_ERROR)
        stream = await llm.add_requests_async(raw_request, token_ids, request.output_len(), request.ignore_eos,
                                              request.temperature, request.top_p, request.top_k,
                                              request.repetition_penalty, mm_contents, stop=request.stop)
        if request.stream:
            return StreamingResponse(chat_completion_stream_generator(stream, request),
                                     media_type="text/event-stream")
        return JSONResponse(content=(await chat_completion_generator(stream, request)).model_dump())

    @app.post("/v1/completions")
    async def create_completion(request: CompletionRequest, raw_request: Request):
        try:
            token_ids = await _in_thread(_encode_prompt, request.prompt)
        except Exception as e:  # noqa: BLE001
            return _error(f"cannot encode prompt: {e}")
        bad = _validate_sampling(request) or _validate(token_ids, request.max_tokens,
                                                       llm.loader.config.get("vocab_size"))
        if bad:
            return _error(bad)
        if llm.failed:
            return _error(f"engine is down: {llm.failed}", HTTPStatus.INTERNAL_SERVER_ERROR)
        stream = await llm.add_requests_async(raw_request, token_ids, request.max_tokens, request.ignore_eos,
                                              request.temperature, request.top_p, request.top_k,
                                              request.repetition_penalty, stop=request.stop)
        if request.stream:
            return StreamingResponse(completion_stream_generator(stream, request), media_type="text/event-stream")
        return JSONResponse(content=(await completion_generator(stream, request)).model_dump())

    @app.post("/tokenize")
    async def tokenize(request: TokenizeRequest):
        toks = llm.encode(request.prompt) if request.prompt is not None else llm.encode(None, True, request.messages)
        return JSONResponse(TokenizeResponse(count=len(toks), max_model_len=llm.model_max_length,
                                             tokens=toks).model_dump())

    @app.post("/detokenize")
    async def detokenize(request: DetokenizeRequest):
        return JSONResponse(DetokenizeResponse(prompt=llm.tokenizer.decode(request.tokens)).model_dump())

    @app.post("/start_profile")
    async def start_profile():
        llm.start_profile()
        return JSONResponse(content={"message": "Profiler started", "success": True})

    @app.post("/stop_profile")
    async def stop_profile():
        llm.stop_profile()
        return JSONResponse(content={"message": "Profiler stopped", "success": True})

    return app


def make_parser() -> argparse.ArgumentParser:
    """CLI surface of the reference (gllm/entrypoints/api_server.py:134-278)."""
    p = argparse.ArgumentParser(description="gllm_b200 OpenAI-compatible server")
    p.add_argument("--host", type=str, default="0.0.0.0")
    p.add_argument("--port", type=int, default=8000)
    p.add_argument("--master-addr", type=str, default="0.0.0.0")
    p.add_argument("--master-port", type=int, default=8001)
    p.add_argument("--zmq-port-base", type=int, default=8002)
    p.add_argument("--model-path", type=str, required=True, help="local HF directory or preset:<name>")
    p.add_argument("--load-format", type=str, choices=["auto", "dummy"], default="auto")
    p.add_argument("--disable-thinking", action="store_true")
    p.add_argument("--model-max-length", type=int, default=None)
    p.add_argument("--use-async-worker", action="store_true")
    p.add_argument("--async-schedule", action=argparse.BooleanOptionalAction, default=True,
                   help="queue the next decode step before the previous step's tokens are back on the host "
                        "(default on; --no-async-schedule for the strictly synchronous loop)")
    p.add_argument("--gpu-memory-util", type=float, default=0.9)
    p.add_argument("--enable-prefix-caching", action="store_true")
    p.add_argument("--page-size", type=int, default=16)
    p.add_argument("--disable-cuda-graph", action="store_true")
    p.add_argument("--max-cuda-graph-bs", type=int, default=512)
    p.add_argument("--pp", type=int, default=1)
    p.add_argument("--tp", type=int, default=1)
    p.add_argument("--disable-ep", action="store_true")
    p.add_argument("--assigned-layers", type=str, default=None, help="e.g. 16,16,17,15")
    p.add_argument("--maxd", type=int, default=2048)
    p.add_argument("--maxp", type=int, default=8192)
    p.add_argument("--minp", type=int, default=32)
    p.add_argument("--iterp", type=int, default=8)
    p.add_argument("--kvthresh", type=float, default=0.05)
    p.add_argument("--schedule-method", type=str, default="chunked_prefill",
                   choices=["split_pd", "chunked_prefill", "token_throttling"])
    p.add_argument("--launch-mode", type=str, default="normal", choices=["normal", "master", "slave"])
    p.add_argument("--ranks", type=str, default=None, help="comma separated global ranks hosted by this node")
    p.add_argument("--mm-processor-min-pixels", type=int, default=None)
    p.add_argument("--mm-processor-max-pixels", type=int, default=None)
    p.add_argument("--tp-mode", type=str, default="fused", choices=["fused", "nccl"])
    return p


def engine_kwargs(args) -> dict:
    return dict(model_path=args.model_path, host=args.host, master_addr=args.master_addr,
                master_port=args.master_port, zmq_port_base=args.zmq_port_base, launch_mode=args.launch_mode,
                worker_ranks=args.ranks, load_format=args.load_format, gpu_memory_util=args.gpu_memory_util,
                page_size=args.page_size, maxd=args.maxd, maxp=args.maxp, minp=args.minp, iterp=args.iterp,
                kvthresh=args.kvthresh, enable_prefix_caching=args.enable_prefix_caching, pp_size=args.pp,
                tp_size=args.tp, use_ep=not args.disable_ep, assigned_layers=args.assigned_layers,
                use_async_worker=args.use_async_worker, async_schedule=args.async_schedule,
                use_thinking=not args.disable_thinking,
                schedule_method=args.schedule_method, disable_cuda_graph=args.disable_cuda_graph,
                max_cuda_graph_bs=args.max_cuda_graph_bs, model_max_length=args.model_max_length,
                mm_processor_min_pixels=args.mm_processor_min_pixels,
                mm_processor_max_pixels=args.mm_processor_max_pixels, tp_mode=args.tp_mode)


async def run_server(app, host: str, port: int):
    import uvicorn
    server = uvicorn.Server(uvicorn.Config(app, port=port, host=host, log_level="warning"))
    task = asyncio.get_running_loop().create_task(server.serve())
    try:
        await task
    except asyncio.CancelledError:
        await server.shutdown()


def main(argv: Optional[List[str]] = None):
    args = make_parser().parse_args(argv)
    from gllm_b200.engine.async_llm_engine import AsyncLLM
    kw = engine_kwargs(args)
    engine = AsyncLLM(kw.pop("model_path"), **kw)
    if args.launch_mode == "slave":
        # a slave node only hosts workers (reference: api_server.py:312-322)
        logger.info("slave node: hosting ranks %s", args.ranks)
        for p in engine.procs:
            p.join()
        return
    app = build_app(engine)
    logger.info("serving on http://%s:%d", args.host, args.port)
    asyncio.run(run_server(app, args.host, args.port))


if __name__ == "__main__":
    main()
