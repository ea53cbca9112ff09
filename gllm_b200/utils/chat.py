"""Chat-history post-processing (reference: `LLM.chat` calls the model's `process_response`, which only the
ChatGLM family defines — gllm/models/chatglm.py:301-325, gllm/llm_engine.py:426).

GLM-3 / GLM-4 answers may hold several `<|assistant|>`-separated segments; a segment whose first line is
non-empty is a tool / code-interpreter call: that first line is the tool name ("metadata"), the rest its payload.
"""
from __future__ import annotations

import ast
import copy
from typing import List, Tuple, Union

GLM_ARCHS = ("ChatGLMModel", "ChatGLMForConditionalGeneration", "GlmForCausalLM", "Glm4ForCausalLM")


def _parse_tool_arguments(payload: str):
    """GLM-3 wraps the call as a fenced snippet `tool_call(key=value, ...)`; return the keyword arguments as a
    dict. Only literals are accepted (`ast.literal_eval` per argument) — the text is model output, never `eval` it."""
    lines = payload.strip().split("\n")
    if len(lines) >= 3 and lines[0].startswith("```"):
        lines = lines[1:-1]
    src = "\n".join(lines).strip()
    try:
        node = ast.parse(src, mode="eval").body
        if isinstance(node, ast.Call):
            return {kw.arg: ast.literal_eval(kw.value) for kw in node.keywords if kw.arg}
        return ast.literal_eval(node)
    except (SyntaxError, ValueError):
        return src


def glm_process_response(output: str, history: List[dict]) -> Tuple[Union[str, dict], List[dict]]:
    """-> (content of the LAST segment, extended copy of history). Plain answers give a string, tool calls a
    dict {"name", "parameters"} (when the system message advertises tools) or {"name", "content"}."""
    history = copy.deepcopy(history)
    has_tools = bool(history) and history[0].get("role") == "system" and "tools" in history[0]
    content: Union[str, dict] = ""
    for segment in output.split("<|assistant|>"):
        metadata, sep, body = segment.partition("\n")
        if not sep:
            metadata, body = "", segment
        if not metadata.strip():
            body = body.strip()
            history.append({"role": "assistant", "metadata": metadata, "content": body})
            content = body
        else:
            history.append({"role": "assistant", "metadata": metadata, "content": body})
            name = metadata.strip()
            content = ({"name": name, "parameters": _parse_tool_arguments(body)} if has_tools
                       else {"name": name, "content": body})
    return content, history


def process_response(architecture: str, output: str, history: List[dict]) -> Tuple[Union[str, dict], List[dict]]:
    """Architecture hook used by `LLM.chat`: GLM models get the segment / tool-call parsing, everything else
    appends one assistant turn."""
    if architecture in GLM_ARCHS:
        return glm_process_response(output, history)
    return output, list(history) + [{"role": "assistant", "content": output}]
