"""OpenAI REST routes (mirrors python/kserve/kserve/protocol/rest/openai/endpoints.py:40-300):
POST {prefix}/v1/completions, {prefix}/v1/chat/completions, GET {prefix}/v1/models; prefix defaults to /openai
and is overridden by KSERVE_OPENAI_ROUTE_PREFIX."""
import os
import time
from typing import AsyncGenerator

from fastapi import APIRouter, FastAPI, Request, Response
from fastapi.responses import JSONResponse, StreamingResponse

from ....errors import ModelNotReady
from .dataplane import OpenAIDataPlane
from .errors import OpenAIError, openai_error_handler
from .types import ChatCompletionRequest, CompletionRequest, ErrorResponse

OPENAI_ROUTE_PREFIX = os.environ.get("KSERVE_OPENAI_ROUTE_PREFIX", "/openai")
if len(OPENAI_ROUTE_PREFIX) > 0 and not OPENAI_ROUTE_PREFIX.startswith("/"):
    OPENAI_ROUTE_PREFIX = f"/{OPENAI_ROUTE_PREFIX}"


class OpenAIEndpoints:
    def __init__(self, dataplane: OpenAIDataPlane):
        self.dataplane = dataplane
        self.start_time = int(time.time())

    async def _respond(self, result):
        if isinstance(result, ErrorResponse):
            return JSONResponse(content=result.model_dump(), status_code=int(result.error.code))
        if isinstance(result, AsyncGenerator):
            return StreamingResponse(result, media_type="text/event-stream")
        # the routes are registered with response_model_exclude_none / response_model_exclude_unset (endpoints.py:262-275):
        # only fields the model set explicitly (and that are not None) reach the client
        return JSONResponse(content=result.model_dump(exclude_none=True, exclude_unset=True))

    async def create_completion(self, request_body: CompletionRequest, raw_request: Request, response: Response):
        model_name = request_body.model
        if not await self.dataplane.model_ready(model_name):
            raise ModelNotReady(model_name)
        completion = await self.dataplane.create_completion(model_name=model_name, request=request_body,
                                                            raw_request=raw_request, headers=raw_request.headers,
                                                            response=response)
        return await self._respond(completion)

    async def create_chat_completion(self, request_body: ChatCompletionRequest, raw_request: Request, response: Response):
        model_name = request_body.model
        if not await self.dataplane.model_ready(model_name):
            raise ModelNotReady(model_name)
        completion = await self.dataplane.create_chat_completion(model_name=model_name, request=request_body,
                                                                 raw_request=raw_request, headers=raw_request.headers,
                                                                 response=response)
        return await self._respond(completion)

    async def models(self):
        """endpoints.py:227-244: {"object": "list", "data": [{"object": "model", "id", "created" (server start), "owned_by": ""}]}"""
        cards = (await self.dataplane.models()).data
        return {"object": "list", "data": [{"id": c.id, "object": "model", "created": self.start_time, "owned_by": ""} for c in cards]}

    async def health(self, model_name: str):
        """GET {prefix}/v1/models/{model_name} (endpoints.py:246-253): 200 when the model is ready, ModelNotReady (503) otherwise"""
        try:
            ready = await self.dataplane.model_ready(model_name)
        except Exception as e:
            raise ModelNotReady(model_name) from e
        if not ready:
            raise ModelNotReady(model_name)


def register_openai_endpoints(app: FastAPI, dataplane: OpenAIDataPlane):
    ep = OpenAIEndpoints(dataplane)
    router = APIRouter(prefix=OPENAI_ROUTE_PREFIX, tags=["openai"])
    router.add_api_route("/v1/completions", ep.create_completion, methods=["POST"], response_model_exclude_none=True)
    router.add_api_route("/v1/chat/completions", ep.create_chat_completion, methods=["POST"], response_model_exclude_none=True)
    router.add_api_route("/v1/models", ep.models, methods=["GET"])
    router.add_api_route("/v1/models/{model_name}", ep.health, methods=["GET"])
    app.include_router(router)
    app.add_exception_handler(OpenAIError, openai_error_handler)
