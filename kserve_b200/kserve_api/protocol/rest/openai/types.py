"""OpenAI request / response models for the completions and chat-completions endpoints.

The reference re-exports these from vLLM (python/kserve/kserve/protocol/rest/openai/types/__init__.py:14-57);
importing vLLM costs tens of seconds and drags its engine in, so the fields the hot path reads are declared
here with the SAME defaults (SURVEY.md §8a' q1: max_tokens=16, stop=[], temperature=None, top_p=None, n=1,
echo=False, stream=False).  Unknown fields are accepted and ignored, as pydantic does for vLLM's models.
"""
from __future__ import annotations

import time
import uuid
from typing import Any, Dict, List, Literal, Optional, Union

from pydantic import BaseModel, ConfigDict, Field


def generate_uuid() -> str:
    return str(uuid.uuid4())


class OpenAIBaseModel(BaseModel):
    model_config = ConfigDict(extra="allow")


class Error(OpenAIBaseModel):
    code: Optional[str] = None
    message: str
    param: Optional[str] = None
    type: str


class ErrorResponse(OpenAIBaseModel):
    error: Error


class UsageInfo(OpenAIBaseModel):
    prompt_tokens: int = 0
    total_tokens: int = 0
    completion_tokens: Optional[int] = 0


class CompletionRequest(OpenAIBaseModel):
    model: Optional[str] = None
    prompt: Optional[Union[List[int], List[List[int]], str, List[str]]] = None
    echo: Optional[bool] = False
    frequency_penalty: Optional[float] = 0.0
    logit_bias: Optional[Dict[str, float]] = None
    logprobs: Optional[int] = None
    max_tokens: Optional[int] = 16
    n: int = 1
    presence_penalty: Optional[float] = 0.0
    seed: Optional[int] = None
    stop: Optional[Union[str, List[str]]] = []
    stream: Optional[bool] = False
    suffix: Optional[str] = None
    temperature: Optional[float] = None
    top_p: Optional[float] = None
    user: Optional[str] = None
    request_id: Optional[str] = None


class CompletionLogProbs(OpenAIBaseModel):
    text_offset: List[int] = Field(default_factory=list)
    token_logprobs: List[Optional[float]] = Field(default_factory=list)
    tokens: List[str] = Field(default_factory=list)
    top_logprobs: List[Optional[Dict[str, float]]] = Field(default_factory=list)


class CompletionChoice(OpenAIBaseModel):
    index: int
    text: str
    logprobs: Optional[CompletionLogProbs] = None
    finish_reason: Optional[str] = None


class Completion(OpenAIBaseModel):
    id: str = Field(default_factory=generate_uuid)
    object: Literal["text_completion"] = "text_completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: Optional[str] = None
    choices: List[CompletionChoice]
    usage: Optional[UsageInfo] = None
    system_fingerprint: Optional[str] = None


class CompletionChunkChoice(CompletionChoice):
    pass


class CompletionChunk(OpenAIBaseModel):
    id: str = Field(default_factory=generate_uuid)
    object: str = "text_completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: Optional[str] = None
    choices: List[CompletionChunkChoice]
    usage: Optional[UsageInfo] = None
    system_fingerprint: Optional[str] = None


class ChatCompletionMessageParam(OpenAIBaseModel):
    """vLLM declares chat messages as TypedDicts: callers index them (`messages[0]["content"]`,
    test_openai_completion.py:324), so the model answers to both attribute and key access."""
    role: str
    content: Optional[Union[str, List[Dict[str, Any]]]] = None
    name: Optional[str] = None

    def __getitem__(self, key: str):
        try:
            return getattr(self, key)
        except AttributeError:
            raise KeyError(key)

    def get(self, key: str, default=None):
        return getattr(self, key, default)


class ChatCompletionRequest(OpenAIBaseModel):
    messages: List[ChatCompletionMessageParam]
    model: Optional[str] = None
    frequency_penalty: Optional[float] = 0.0
    logit_bias: Optional[Dict[str, float]] = None
    logprobs: Optional[bool] = False
    top_logprobs: Optional[int] = None
    max_tokens: Optional[int] = None
    n: Optional[int] = 1
    presence_penalty: Optional[float] = 0.0
    seed: Optional[int] = None
    stop: Optional[Union[str, List[str]]] = []
    stream: Optional[bool] = False
    temperature: Optional[float] = None
    top_p: Optional[float] = None
    tools: Optional[List[Dict[str, Any]]] = None
    user: Optional[str] = None
    chat_template: Optional[str] = None
    add_generation_prompt: bool = True
    continue_final_message: bool = False
    documents: Optional[List[Dict[str, str]]] = None
    chat_template_kwargs: Optional[Dict[str, Any]] = None
    request_id: Optional[str] = None


class ChatMessage(OpenAIBaseModel):
    role: str
    content: Optional[str] = None
    tool_calls: List[Any] = Field(default_factory=list)     # vLLM's ChatMessage carries an (empty) list: fixtures/openai/chat_completion.json


class ChatCompletionLogProb(OpenAIBaseModel):
    token: str
    logprob: float = -9999.0
    bytes: Optional[List[int]] = None


class ChatCompletionLogProbsContent(ChatCompletionLogProb):
    top_logprobs: List[ChatCompletionLogProb] = Field(default_factory=list)


class ChatCompletionLogProbs(OpenAIBaseModel):
    content: Optional[List[ChatCompletionLogProbsContent]] = None


class ChatCompletionChoice(OpenAIBaseModel):
    index: int
    message: ChatMessage
    logprobs: Optional[ChatCompletionLogProbs] = None
    finish_reason: Optional[str] = "stop"


class ChatCompletion(OpenAIBaseModel):
    id: str
    object: Literal["chat.completion"] = "chat.completion"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: Optional[str] = None
    choices: List[ChatCompletionChoice]
    usage: Optional[UsageInfo] = None
    system_fingerprint: Optional[str] = None


class ChoiceDelta(OpenAIBaseModel):
    role: Optional[str] = None
    content: Optional[str] = None


class ChunkChoice(OpenAIBaseModel):
    index: int
    delta: ChoiceDelta
    logprobs: Optional[ChatCompletionLogProbs] = None
    finish_reason: Optional[str] = None


class ChatCompletionChunk(OpenAIBaseModel):
    id: str
    object: Literal["chat.completion.chunk"] = "chat.completion.chunk"
    created: int = Field(default_factory=lambda: int(time.time()))
    model: Optional[str] = None
    choices: List[ChunkChoice]
    usage: Optional[UsageInfo] = None
    system_fingerprint: Optional[str] = None


class ModelCard(OpenAIBaseModel):
    id: str
    object: str = "model"
    created: int = Field(default_factory=lambda: int(time.time()))
    owned_by: str = "kserve"


class ModelList(OpenAIBaseModel):
    object: str = "list"
    data: List[ModelCard] = Field(default_factory=list)
