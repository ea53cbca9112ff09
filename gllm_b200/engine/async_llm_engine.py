"""asyncio front-end for the HTTP server (reference: gllm/async_llm_engine.py:11-109).

One `AsyncStream` per request (an asyncio.Queue of text deltas); a single background task ticks the
engine (`LLM.schedule`) in a worker thread so the event loop stays responsive, detokenises new
tokens incrementally, and aborts requests whose client disconnected.
"""
from __future__ import annotations

import asyncio
import time
from typing import Dict, List, Optional

from gllm_b200.engine.llm_engine import LLM
from gllm_b200.utils.logging import logger


class Histogram:
    """Cumulative-bucket latency histogram in the Prometheus exposition layout (seconds)."""

    def __init__(self, bounds):
        self.bounds = tuple(bounds)
        self.counts = [0] * (len(self.bounds) + 1)
        self.sum = 0.0
        self.count = 0

    def observe(self, v: float):
        self.sum += v
        self.count += 1
        for i, b in enumerate(self.bounds):
            if v <= b:
                self.counts[i] += 1
                return
        self.counts[-1] += 1

    def lines(self, name: str) -> List[str]:
        out, acc = [f"# TYPE {name} histogram"], 0
        for b, c in zip(self.bounds, self.counts):
            acc += c
            out.append(f'{name}_bucket{{le="{b:g}"}} {acc}')
        out.append(f'{name}_bucket{{le="+Inf"}} {self.count}')
        out.append(f"{name}_sum {self.sum:.6f}")
        out.append(f"{name}_count {self.count}")
        return out


TTFT_BUCKETS = (0.005, 0.01, 0.02, 0.04, 0.06, 0.08, 0.1, 0.25, 0.5, 0.75, 1.0, 2.5, 5.0, 7.5, 10.0, 30.0)
TPOT_BUCKETS = (0.002, 0.004, 0.006, 0.008, 0.01, 0.015, 0.02, 0.03, 0.04, 0.05, 0.075, 0.1, 0.2, 0.5, 1.0)
E2E_BUCKETS = (0.1, 0.25, 0.5, 1.0, 2.5, 5.0, 10.0, 20.0, 40.0, 60.0, 120.0, 300.0)


class AsyncStream:
    def __init__(self, raw_request=None, stop=None):
        self.stop = [x for x in ([stop] if isinstance(stop, str) else list(stop or [])) if x]
        self._held = ""
        self.stop_hit = False
        self._queue: asyncio.Queue = asyncio.Queue()
        self._finished = False
        self._raw_request = raw_request
        self.prompt_tokens = 0
        self.completion_tokens = 0
        self.finish_reason: Optional[str] = None
        self.created = time.time()
        self.first_token_time: Optional[float] = None
        self.seq_id = -1
        self.aborted = False     # an abort id was sent for this request (front-end side state only)

    def put(self, item: str):
        """Queue a text delta. With stop strings (OpenAI `stop`; not honoured by the reference) the text that could
        still turn into a stop string is held back; on a hit the text before it is delivered, the stream ends with
        finish_reason "stop" and `stop_hit` tells the engine to abort the request."""
        if self._finished:
            return
        if not self.stop:
            self._queue.put_nowait(item)
            return
        buf = self._held + item
        cut = min((i for i in (buf.find(s) for s in self.stop) if i >= 0), default=-1)
        if cut >= 0:
            self._held = ""
            if cut:
                self._queue.put_nowait(buf[:cut])
            self.stop_hit = True
            self.finish("stop")
            return
        hold = 0      # longest suffix of buf that is a proper prefix of some stop string
        for s in self.stop:
            for k in range(min(len(s) - 1, len(buf)), hold, -1):
                if buf.endswith(s[:k]):
                    hold = k
                    break
        if len(buf) > hold:
            self._queue.put_nowait(buf[:len(buf) - hold])
        self._held = buf[len(buf) - hold:] if hold else ""

    def finish(self, reason: str = "stop"):
        if not self._finished:
            if self._held:                     # generation ended while a possible stop prefix was held back
                self._queue.put_nowait(self._held)
                self._held = ""
            self.finish_reason = reason
            self._queue.put_nowait(StopAsyncIteration())
            self._finished = True

    @property
    def finished(self) -> bool:
        return self._finished

    def __aiter__(self):
        return self

    async def __anext__(self):
        item = await self._queue.get()
        if isinstance(item, Exception):
            raise item
        return item

    async def is_disconnected(self) -> bool:
        if self._raw_request is None:
            return False
        try:
            return await self._raw_request.is_disconnected()
        except Exception:  # noqa: BLE001
            return False


class AsyncLLM(LLM):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.async_streams: Dict[int, AsyncStream] = {}
        self._task: Optional[asyncio.Task] = None
        self._pending_tokens: List = []
        self.failed: Optional[str] = None      # set when the engine loop died: the server answers 500 from then on
        self.metrics = {"requests_total": 0, "requests_finished": 0, "requests_aborted": 0,
                        "prompt_tokens_total": 0, "generation_tokens_total": 0, "ttft_sum": 0.0, "ttft_count": 0}
        self.hist = {"ttft": Histogram(TTFT_BUCKETS), "tpot": Histogram(TPOT_BUCKETS),
                     "e2e": Histogram(E2E_BUCKETS)}

    async def add_requests_async(self, raw_request, token_ids: List[int], output_len=None, ignore_eos=False,
                                 temperature=None, top_p=None, top_k=None, repetition_penalty=None,
                                 mm_contents=None, stop=None) -> AsyncStream:
        seq = self.allocate_seq(token_ids, output_len, ignore_eos, temperature, top_p, top_k, repetition_penalty,
                                mm_contents)
        stream = AsyncStream(raw_request, stop)
        stream.prompt_tokens = len(token_ids)
        stream.seq_id = seq.seq_id
        self.async_streams[seq.seq_id] = stream
        self.metrics["requests_total"] += 1
        self.metrics["prompt_tokens_total"] += len(token_ids)
        self.add_requests([seq])
        if self._task is None and self.failed is None:
            self.start_schedule_engine()
        return stream

    def abort_stream(self, stream: AsyncStream):
        """Client went away: abort the request so its KV pages are freed (reference:
        async_llm_engine.py:93-97 polls `is_disconnected` from the engine task; here the request's own
        task reports it, which also works under ASGI test transports)."""
        # Only the abort id is sent: the Sequence object is shared with the in-proc scheduler, which must be the
        # one to take it out of its queues before flagging it (a flag set from here while a prefill chunk was in
        # flight left a freed sequence at the head of the prefill queue and took the engine down).
        if not stream.aborted and not stream.finished:
            stream.aborted = True
            self.abort([stream.seq_id])
            self.metrics["requests_aborted"] += 1

    async def collect(self, stream: AsyncStream) -> str:
        """Drain a stream to a string, aborting the request if the client disconnects meanwhile."""
        text = ""
        while True:
            try:
                text += await asyncio.wait_for(stream.__anext__(), timeout=0.5)
            except StopAsyncIteration:
                return text
            except asyncio.TimeoutError:
                if await stream.is_disconnected():
                    self.abort_stream(stream)
                    return text

    def _on_token(self, seq, tok):
        self._pending_tokens.append(seq)

    def _tick(self) -> bool:
        return self.schedule(self._on_token)

    def _deliver(self):
        seen = set()
        for seq in self._pending_tokens:
            if id(seq) in seen:
                continue
            seen.add(id(seq))
            st = self.async_streams.get(seq.seq_id)
            if st is None:
                continue
            if st.first_token_time is None:
                st.first_token_time = time.time()
                self.metrics["ttft_sum"] += st.first_token_time - st.created
                self.metrics["ttft_count"] += 1
                self.hist["ttft"].observe(st.first_token_time - st.created)
            st.completion_tokens = seq.num_output_tokens
            if self.tokenizer is not None:
                delta = seq.detokenize_inc(self.tokenizer)
                if delta:
                    st.put(delta)
            else:
                st.put(" ".join(str(t) for t in seq.token_ids[seq.cur_length:seq.known_len]) + " ")
                seq.cur_length = seq.known_len
            if st.stop_hit and not st.aborted:      # a stop string completed: stop generating for this request
                st.aborted = True
                self.abort([seq.seq_id])
        self._pending_tokens = []
        for seq in self.finished:
            st = self.async_streams.pop(seq.seq_id, None)
            if st is not None:
                st.completion_tokens = seq.num_output_tokens
                self.metrics["generation_tokens_total"] += seq.num_output_tokens
                self.metrics["requests_finished"] += 1
                now = time.time()
                self.hist["e2e"].observe(now - st.created)
                if st.first_token_time is not None and seq.num_output_tokens > 1:
                    self.hist["tpot"].observe((now - st.first_token_time) / (seq.num_output_tokens - 1))
                reason = "length" if seq.num_output_tokens >= seq.output_len else "stop"
                if st.stop_hit:
                    reason = "stop"
                elif st.aborted or seq.is_abort:
                    reason = "abort"
                st.finish(reason)
        self.finished = []

    async def _loop(self):
        loop = asyncio.get_running_loop()
        idle = 0
        while True:
            did = await loop.run_in_executor(None, self._tick)
            self._deliver()
            if did or self.running_maps or self.wait_lists:
                idle = 0
                await asyncio.sleep(0)
            else:
                idle += 1
                await asyncio.sleep(0.001 if idle < 1000 else 0.01)

    def start_schedule_engine(self):
        self._task = asyncio.get_event_loop().create_task(self._loop())

        def _done(task):
            try:
                task.result()
            except asyncio.CancelledError:
                logger.info("engine loop cancelled")
            except Exception as e:  # noqa: BLE001
                logger.error("engine background task failed: %r", e, exc_info=e)
                # fail-stop: release every waiting client instead of hanging them
                self.failed = repr(e)
                for st in list(self.async_streams.values()):
                    st._queue.put_nowait(RuntimeError(f"engine failure: {e!r}"))
                self.async_streams.clear()
                self._task = None
        self._task.add_done_callback(_done)


# name used by the reference (gllm/async_llm_engine.py:56)
PipeAsyncLLM = AsyncLLM
