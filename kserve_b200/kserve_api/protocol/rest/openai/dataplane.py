"""OpenAIDataPlane (mirrors python/kserve/kserve/protocol/rest/openai/dataplane.py:41-177)."""
from typing import AsyncGenerator, Union

from fastapi import Request, Response

from ....errors import InvalidInput
from ...dataplane import DataPlane
from .openai_model import OpenAIGenerativeModel, OpenAIModel
from .types import (ChatCompletion, ChatCompletionRequest, Completion, CompletionRequest, ErrorResponse, ModelCard,
                    ModelList)


class OpenAIDataPlane(DataPlane):
    async def create_completion(self, model_name: str, request: CompletionRequest, raw_request: Request,
                                headers, response: Response
                                ) -> Union[AsyncGenerator[str, None], Completion, ErrorResponse]:
        model = await self.get_model(model_name)
        if not isinstance(model, OpenAIGenerativeModel):
            raise InvalidInput(f"Model {model_name} does not support Completions API")
        context = {"headers": dict(headers), "response": response}
        return await model.create_completion(request=request, raw_request=raw_request, context=context)

    async def create_chat_completion(self, model_name: str, request: ChatCompletionRequest, raw_request: Request,
                                     headers, response: Response
                                     ) -> Union[AsyncGenerator[str, None], ChatCompletion, ErrorResponse]:
        model = await self.get_model(model_name)
        if not isinstance(model, OpenAIGenerativeModel):
            raise InvalidInput(f"Model {model_name} does not support Chat Completion API")
        context = {"headers": dict(headers), "response": response}
        return await model.create_chat_completion(request=request, raw_request=raw_request, context=context)

    async def models(self) -> ModelList:
        return ModelList(data=[ModelCard(id=name) for name, m in self.model_registry.get_models().items()
                               if isinstance(m, OpenAIModel)])
