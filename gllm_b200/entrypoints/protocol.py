"""OpenAI-compatible request / response schemas (pydantic v2).

Written against the public OpenAI API shape; field coverage follows what the reference accepts
(gllm/entrypoints/protocol.py:165-716) plus a filled `usage`. Unknown request fields are ignored
rather than rejected so stock OpenAI clients work unchanged.
"""
from __future__ import annotations

import time
import uuid
from typing import Any, Dict, List, Literal, Optional, Union

from pydantic import BaseModel, ConfigDict, Field


def _rid(prefix: str) -> str:
    return f"{prefix}-{uuid.uuid4().hex}"


class OpenAIBase(BaseModel):
    model_config = ConfigDict(extra="allow")


class ErrorResponse(OpenAIBase):
    object: str = "error"
    message: str
    type: str
    param: Optional[str] = None
    code: int


class ModelPermission(OpenAIBase):
    id: str = Field(default_factory=lambda: _rid("modelperm"))
    object: str = "model_permission"
    created: int = Field(default_factory=lambda: int(time.time()))
    allow_create_engine: bool = False
    allow_sampling: bool = True
    allow_logprobs: bool = True
    allow_search_indices: bool = False
    allow_view: bool = True
    allow_fine_tuning: bool = False
    organization: str = "*"
    group: Optional[str] = None
    is_blocking: bool = False


class ModelCard(OpenAIBase):
    id: str
    object: str = "model"
    created: int = Field(default_factory=lambda: int(time.time()))
    owned_by: str = "gllm_b200"
    root: Optional[str] = None
    parent: Optional[str] = None
    max_model_len: Optional[int] = None
    permission: List[ModelPermission] = Field(default_factory=list)


class ModelList(OpenAIBase):
    object: str = "list"
    data: List[ModelCard] = Field(default_factory=list)


class UsageInfo(OpenAIBase):
    prompt_tokens: int = 0
    total_tokens: int = 0
    completion_tokens: Optional[int] = 0


class StreamOptions(OpenAIBase):
    include_usage: Optional[bool] = True
    continuous_usage_stats: Optional[bool] = False


class _SamplingMixin(OpenAIBase):
    temperature: Optional[float] = None
    top_p: Optional[float] = None
    top_k: Optional[int] = None
    repetition_penalty: Optional[float] = None
    ignore_eos: bool = False
    stream: Optional[bool] = False
    stream_options: Optional[StreamOptions] = None
    n: Optional[int] = 1
    seed: Optional[int] = None
    stop: Optional[Union[str, List[str]]] = None
    frequency_penalty: Optional[float] = 0.0
    presence_penalty: Optional[float] = 0.0
    logit_bias: Optional[Dict[str, float]] = None
    user: Optional[str] = None


class ChatCompletionRequest(_SamplingMixin):
    messages: List[Dict[str, Any]]
    model: Optional[str] = None
    max_tokens: Optional[int] = None
    max_completion_tokens: Optional[int] = None
    logprobs: Optional[bool] = False
    top_logprobs: Optional[int] = 0
    tools: Optional[List[Dict[str, Any]]] = None
    tool_choice: Optional[Union[str, Dict[str, Any]]] = None
    response_format: Optional[Dict[str, Any]] = None
    chat_template_kwargs: Optional[Dict[str, Any]] = None

    def output_len(self) -> Optional[int]:
        return self.max_completion_tokens if self.max_completion_tokens is not None else self.max_tokens


class CompletionRequest(_SamplingMixin):
    model: Optional[str] = None
    prompt: Union[str, List[str], List[int], List[List[int]]]
    max_tokens: Optional[int] = 16
    echo: Optional[bool] = False
    logprobs: Optional[int] = None
    suffix: Optional[str] = None
    best_of: Optional[int] = None


class ChatMessage(OpenAIBase):
    role: str
    content: Optional[str] = None
    reasoning_content: Optional[str] = None
    tool_calls: List[Dict[str, Any]] = Field(default_factory=list)


class ChatCompletionResponseChoice(OpenAIBase):
    index: int
    message: ChatMessage
    logprobs: Optional[Any] = None
    finish_reason: Optional[str] = "stop"


class ChatCompletionResponse(OpenAIBase):
    id: str = Field(default_factory=lambda: _rid("chatcmpl"))
    object: Literal["chat.completion"] = "chat.completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: Optional[str] = None
    choices: List[ChatCompletionResponseChoice]
    usage: UsageInfo


class DeltaMessage(OpenAIBase):
    role: Optional[str] = None
    content: Optional[str] = None
    reasoning_content: Optional[str] = None
    tool_calls: List[Dict[str, Any]] = Field(default_factory=list)


class ChatCompletionResponseStreamChoice(OpenAIBase):
    index: int
    delta: DeltaMessage
    logprobs: Optional[Any] = None
    finish_reason: Optional[str] = None


class ChatCompletionStreamResponse(OpenAIBase):
    id: str = Field(default_factory=lambda: _rid("chatcmpl"))
    object: Literal["chat.completion.chunk"] = "chat.completion.chunk"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: Optional[str] = None
    choices: List[ChatCompletionResponseStreamChoice]
    usage: Optional[UsageInfo] = None


class CompletionResponseChoice(OpenAIBase):
    index: int
    text: str
    logprobs: Optional[Any] = None
    finish_reason: Optional[str] = "stop"


class CompletionResponse(OpenAIBase):
    id: str = Field(default_factory=lambda: _rid("cmpl"))
    object: str = "text_completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: Optional[str] = None
    choices: List[CompletionResponseChoice]
    usage: UsageInfo


class CompletionResponseStreamChoice(OpenAIBase):
    index: int
    text: str
    logprobs: Optional[Any] = None
    finish_reason: Optional[str] = None


class CompletionStreamResponse(OpenAIBase):
    id: str = Field(default_factory=lambda: _rid("cmpl"))
    object: str = "text_completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: Optional[str] = None
    choices: List[CompletionResponseStreamChoice]
    usage: Optional[UsageInfo] = None


class TokenizeRequest(OpenAIBase):
    model: Optional[str] = None
    prompt: Optional[str] = None
    messages: Optional[List[Dict[str, Any]]] = None
    add_special_tokens: bool = True


class TokenizeResponse(OpenAIBase):
    count: int
    max_model_len: int
    tokens: List[int]


class DetokenizeRequest(OpenAIBase):
    model: Optional[str] = None
    tokens: List[int]


class DetokenizeResponse(OpenAIBase):
    prompt: str
