"""Maps /v1/chat/completions onto /v1/completions (mirrors
python/kserve/kserve/protocol/rest/openai/openai_chat_adapter_model.py:48-244): a subclass supplies
apply_chat_template() and create_completion(); chat responses and SSE chunks are re-wrapped here."""
from abc import abstractmethod
from typing import Any, AsyncGenerator, Dict, Optional, Union, cast

from fastapi import Request

from ....errors import InvalidInput
from .openai_model import AsyncMappingIterator, ChatPrompt, OpenAIGenerativeModel
from .types import (ChatCompletion, ChatCompletionChoice, ChatCompletionChunk, ChatCompletionLogProb,
                    ChatCompletionLogProbs, ChatCompletionLogProbsContent, ChatCompletionRequest, ChatMessage,
                    ChoiceDelta, ChunkChoice, Completion, CompletionChoice, CompletionChunk, CompletionChunkChoice,
                    CompletionLogProbs, CompletionRequest, ErrorResponse)


class OpenAIChatAdapterModel(OpenAIGenerativeModel):
    @abstractmethod
    def apply_chat_template(self, request: ChatCompletionRequest) -> ChatPrompt:
        pass

    @classmethod
    def chat_completion_params_to_completion_params(cls, request: ChatCompletionRequest, prompt: str) -> CompletionRequest:
        # :66-86 — note logprobs <- top_logprobs and max_tokens passed through (None => model max length)
        return CompletionRequest(
            prompt=prompt, model=request.model, frequency_penalty=request.frequency_penalty,
            logit_bias=request.logit_bias, max_tokens=request.max_tokens, n=request.n,
            presence_penalty=request.presence_penalty, seed=request.seed, stop=request.stop, stream=request.stream,
            temperature=request.temperature, top_p=request.top_p, user=request.user, logprobs=request.top_logprobs,
            request_id=request.request_id)

    @classmethod
    def to_choice_logprobs(cls, lp: CompletionLogProbs) -> ChatCompletionLogProbs:
        """:88-113 — completions-style parallel lists (tokens / token_logprobs / top_logprobs dicts) become the chat
        API's per-token records with the UTF-8 bytes of each token (pinned by fixtures/openai/chat_completion*.json)."""
        def rec(token: str, logprob) -> dict:
            return dict(token=token, bytes=list(token.encode("utf8")), logprob=logprob)
        content = []
        for i, token in enumerate(lp.tokens):
            tops = [ChatCompletionLogProb(**rec(t, v)) for t, v in (lp.top_logprobs[i] or {}).items()]
            content.append(ChatCompletionLogProbsContent(**rec(token, lp.token_logprobs[i]), top_logprobs=tops))
        return ChatCompletionLogProbs(content=content)

    @classmethod
    def to_chat_completion_choice(cls, choice: CompletionChoice, role: str) -> ChatCompletionChoice:
        lp = cls.to_choice_logprobs(choice.logprobs) if choice.logprobs is not None else None
        return ChatCompletionChoice(index=0, finish_reason=choice.finish_reason, logprobs=lp,
                                    message=ChatMessage(content=choice.text, role=role))

    @classmethod
    def to_chat_completion_chunk_choice(cls, choice: CompletionChunkChoice, role: str) -> ChunkChoice:
        lp = cls.to_choice_logprobs(choice.logprobs) if choice.logprobs is not None else None
        return ChunkChoice(delta=ChoiceDelta(content=choice.text, role=role), index=0,
                           finish_reason=choice.finish_reason, logprobs=lp)

    @classmethod
    def completion_to_chat_completion(cls, completion: Completion, role: str) -> ChatCompletion:
        choices = [cls.to_chat_completion_choice(completion.choices[0], role)] if completion.choices else []
        return ChatCompletion(id=completion.id, choices=choices, created=completion.created, model=completion.model,
                              object="chat.completion", system_fingerprint=completion.system_fingerprint, usage=completion.usage)

    @classmethod
    def completion_to_chat_completion_chunk(cls, completion: CompletionChunk, role: str) -> ChatCompletionChunk:
        choices = [cls.to_chat_completion_chunk_choice(completion.choices[0], role)] if completion.choices else []
        return ChatCompletionChunk(id=completion.id, choices=choices, created=completion.created, model=completion.model,
                                   object="chat.completion.chunk", system_fingerprint=completion.system_fingerprint)

    async def create_chat_completion(self, request: ChatCompletionRequest, raw_request: Optional[Request] = None,
                                     context: Optional[Dict[str, Any]] = None
                                     ) -> Union[AsyncGenerator[str, None], ChatCompletion, ErrorResponse]:
        if request.n != 1:
            raise InvalidInput("n != 1 is not supported")
        chat_prompt = self.apply_chat_template(request)
        params = self.chat_completion_params_to_completion_params(request, chat_prompt.prompt)
        if not request.stream:
            completion = cast(Completion, await self.create_completion(params, raw_request))
            return self.completion_to_chat_completion(completion, chat_prompt.response_role)
        iterator = await self.create_completion(params, raw_request)

        def mapper(s: str):
            chunk = s.removeprefix("data: ")
            if chunk == "[DONE]\n\n":
                return None
            return self.completion_to_chat_completion_chunk(CompletionChunk.model_validate_json(chunk),
                                                            chat_prompt.response_role)

        mapped = AsyncMappingIterator(iterator=iterator, mapper=mapper)

        async def stream_results() -> AsyncGenerator[str, None]:
            async for chunk in mapped:
                yield f"data: {chunk.model_dump_json()}\n\n"
            yield "data: [DONE]\n\n"

        return stream_results()
