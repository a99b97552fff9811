"""Known answers for the chat adapter: the completions -> chat-completions re-wrap of
python/kserve/kserve/protocol/rest/openai/openai_chat_adapter_model.py:66-244, against the values the reference's
own fixtures pin (python/kserve/test/fixtures/openai/{completion,chat_completion,completion_partial,
chat_completion_chunk,*_create_params}.json, exercised by test_openai_completion.py:196-335; abbreviated here to the
first two tokens)."""
import asyncio
import json

from kserve_b200.kserve_api.protocol.rest.openai.openai_chat_adapter_model import OpenAIChatAdapterModel
from kserve_b200.kserve_api.protocol.rest.openai.openai_model import ChatPrompt
from kserve_b200.kserve_api.protocol.rest.openai.types import (ChatCompletion, ChatCompletionChunk, ChatCompletionRequest,
                                                               Completion, CompletionRequest)

COMPLETION = {
    "id": "61ccd360-5fca-446d-ad04-8ea32699764b", "object": "text_completion", "created": 1709934260, "model": "gpt-3.5",
    "choices": [{"text": "\n\nAI", "index": 0, "finish_reason": "length",
                 "logprobs": {"tokens": ["\n\n", "AI"], "token_logprobs": [-0.27425793, -0.016520381],
                              "top_logprobs": [{"\n\n": -0.27425793}, {"AI": -0.016520381}], "text_offset": [11, 13]}}],
    "usage": {"prompt_tokens": 4, "completion_tokens": 7, "total_tokens": 11}, "system_fingerprint": "fp_4f0b692a78"}
CHAT_COMPLETION = {
    "id": "61ccd360-5fca-446d-ad04-8ea32699764b", "object": "chat.completion", "created": 1709934260, "model": "gpt-3.5",
    "choices": [{"finish_reason": "length", "index": 0,
                 "logprobs": {"content": [
                     {"token": "\n\n", "bytes": [10, 10], "logprob": -0.27425793,
                      "top_logprobs": [{"token": "\n\n", "bytes": [10, 10], "logprob": -0.27425793}]},
                     {"token": "AI", "bytes": [65, 73], "logprob": -0.016520381,
                      "top_logprobs": [{"token": "AI", "bytes": [65, 73], "logprob": -0.016520381}]}]},
                 "message": {"content": "\n\nAI", "role": "assistant", "tool_calls": []}}],
    "system_fingerprint": "fp_4f0b692a78", "usage": {"completion_tokens": 7, "prompt_tokens": 4, "total_tokens": 11}}
PARTIAL = {
    "id": "c0d5dd7e-9bff-4a68-8cd2-b743612385ac", "object": "text_completion", "created": 1709996697, "model": "gpt-3.5-turbo",
    "choices": [{"text": " intelligence", "index": 0, "finish_reason": "stop",
                 "logprobs": {"tokens": [" intelligence"], "token_logprobs": [-0.00023035755],
                              "top_logprobs": [{" intelligence": -0.00023035755}], "text_offset": [37]}}],
    "system_fingerprint": "fp_4f0b692a78", "usage": {"prompt_tokens": 7, "completion_tokens": 1, "total_tokens": 8}}
_B = [32, 105, 110, 116, 101, 108, 108, 105, 103, 101, 110, 99, 101]
CHUNK = {
    "id": "c0d5dd7e-9bff-4a68-8cd2-b743612385ac", "object": "chat.completion.chunk", "created": 1709996697,
    "model": "gpt-3.5-turbo", "system_fingerprint": "fp_4f0b692a78",
    "choices": [{"index": 0, "delta": {"content": " intelligence", "role": "assistant"}, "finish_reason": "stop",
                 "logprobs": {"content": [{"token": " intelligence", "logprob": -0.00023035755, "bytes": _B,
                                           "top_logprobs": [{"token": " intelligence", "logprob": -0.00023035755, "bytes": _B}]}]}}]}


class Dummy(OpenAIChatAdapterModel):
    """same shape as the reference's DummyModel (test_openai_completion.py:72-103)"""
    def __init__(self, full, partial, n=5):
        self.full, self.partial, self.n = full, partial, n

    async def create_completion(self, request, raw_request=None, context=None):
        if not request.stream:
            return self.full

        async def gen():
            for _ in range(self.n):
                yield f"data: {self.partial.model_dump_json()}\n\n"
            yield "data: [DONE]\n\n"
        return gen()

    def apply_chat_template(self, request):
        return ChatPrompt(prompt="hello")


def test_completion_to_chat_completion_with_logprobs_and_tool_calls():
    got = OpenAIChatAdapterModel.completion_to_chat_completion(Completion.model_validate(COMPLETION), "assistant")
    assert got.model_dump_json() == ChatCompletion.model_validate(CHAT_COMPLETION).model_dump_json()
    got = OpenAIChatAdapterModel.completion_to_chat_completion_chunk(Completion.model_validate(PARTIAL), "assistant")
    assert got.model_dump_json(indent=2) == ChatCompletionChunk.model_validate(CHUNK).model_dump_json(indent=2)


def test_chat_params_to_completion_params():
    chat = ChatCompletionRequest.model_validate({"model": "gpt-3.5-turbo", "messages": [{"role": "user", "content": "What is AI?"}],
                                                 "max_tokens": 7, "logprobs": True, "top_logprobs": 1})
    want = CompletionRequest.model_validate({"model": "gpt-3.5-turbo", "prompt": "What is AI?", "max_tokens": 7, "logprobs": 1})
    got = OpenAIChatAdapterModel.chat_completion_params_to_completion_params(chat, prompt=chat.messages[0]["content"])
    assert got == want.model_copy(update={"request_id": got.request_id})       # logprobs <- top_logprobs (:66-86)


def test_create_chat_completion_plain_and_streaming():
    m = Dummy(Completion.model_validate(COMPLETION), Completion.model_validate(PARTIAL))
    req = ChatCompletionRequest.model_validate({"model": "gpt-3.5-turbo", "messages": [{"role": "user", "content": "What is AI?"}]})
    c = asyncio.run(m.create_chat_completion(req))
    assert c.model_dump_json(indent=2) == ChatCompletion.model_validate(CHAT_COMPLETION).model_dump_json(indent=2)
    req.stream = True

    async def drain():
        out = []
        async for s in await m.create_chat_completion(req):
            out.append(s)
        return out
    chunks = asyncio.run(drain())
    want = "data: " + ChatCompletionChunk.model_validate(CHUNK).model_dump_json() + "\n\n"   # no indent, two newlines
    assert chunks == [want] * 5 + ["data: [DONE]\n\n"]
    assert json.loads(chunks[0][6:])["choices"][0]["logprobs"]["content"][0]["bytes"] == _B
