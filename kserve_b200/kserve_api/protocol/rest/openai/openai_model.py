"""OpenAIModel / OpenAIGenerativeModel plug-in interfaces (mirrors
python/kserve/kserve/protocol/rest/openai/openai_model.py:36-107)."""
from abc import abstractmethod
from typing import Any, AsyncGenerator, Callable, Dict, Optional, Union

from fastapi import Request
from pydantic import BaseModel

from ....model import BaseKServeModel
from .types import ChatCompletion, ChatCompletionRequest, Completion, CompletionRequest, ErrorResponse


class ChatPrompt(BaseModel):
    response_role: str = "assistant"
    prompt: str


class OpenAIModel(BaseKServeModel):
    def __init__(self, name: str):
        super().__init__(name)
        self.ready = True  # openai_model.py:50-52: load() is not part of this interface yet


class OpenAIGenerativeModel(OpenAIModel):
    @abstractmethod
    async def create_completion(self, request: CompletionRequest, raw_request: Optional[Request] = None,
                                context: Optional[Dict[str, Any]] = None
                                ) -> Union[AsyncGenerator[str, None], Completion, ErrorResponse]:
        pass

    @abstractmethod
    async def create_chat_completion(self, request: ChatCompletionRequest, raw_request: Optional[Request] = None,
                                     context: Optional[Dict[str, Any]] = None
                                     ) -> Union[AsyncGenerator[str, None], ChatCompletion, ErrorResponse]:
        pass


class AsyncMappingIterator:
    """Maps an async iterator through `mapper`, dropping None results (openai_model.py:110-133)."""

    def __init__(self, iterator, mapper: Callable = lambda x: x, skip_none: bool = True, close: Optional[Callable] = None):
        self.iterator, self.mapper, self.skip_none, self.close = iterator, mapper, skip_none, close

    def __aiter__(self):
        return self

    async def __anext__(self):
        while True:
            try:
                item = await self.iterator.__anext__()
            except StopAsyncIteration:
                if self.close:
                    self.close()
                raise
            out = self.mapper(item)
            if out is None and self.skip_none:
                continue
            return out
