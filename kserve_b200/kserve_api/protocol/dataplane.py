"""DataPlane: model lookup / readiness / decode / infer / encode (mirrors
python/kserve/kserve/protocol/dataplane.py:59-507 for the REST legs; CloudEvents are out of scope)."""
from __future__ import annotations

import json
import time
from typing import Dict, Optional, Tuple, Union

from ..errors import InvalidInput, ModelNotFound
from ..model import BaseKServeModel, InferenceModel
from ..model_repository import ModelRepository
from .infer_type import InferRequest, InferResponse

SERVER_NAME = "kserve"   # what GET /v2 reports in the reference (dataplane.py metadata(), test_dataplane.py:124-135)
SERVER_VERSION = "0.1.0"


class DataPlane:
    def __init__(self, model_registry: ModelRepository):
        self._model_registry = model_registry
        self._server_name = SERVER_NAME
        self._server_version = SERVER_VERSION
        self._start = time.time()

    @property
    def model_registry(self):
        return self._model_registry

    def get_model_from_registry(self, name: str) -> BaseKServeModel:
        model = self._model_registry.get_model(name)
        if model is None:
            raise ModelNotFound(name)
        return model

    async def get_model(self, name: str) -> BaseKServeModel:
        """dataplane.py:111-126 — loads the model on first use if it is registered but not ready."""
        model = self.get_model_from_registry(name)
        if not await self._model_registry.is_model_ready(name):
            model.load()
        return model

    @staticmethod
    async def live() -> Dict[str, str]:
        return {"status": "alive"}

    async def ready(self) -> bool:
        """dataplane.py:247-279: the server readiness probe answers True whatever the readiness of the registered models
        (those have their own probes); only a transformer in front of a remote predictor forwards the question."""
        return True

    def metadata(self) -> Dict:
        return {"name": self._server_name, "version": self._server_version, "extensions": ["model_repository_extension"]}

    async def model_metadata(self, model_name: str) -> Dict:
        model = self.get_model_from_registry(model_name)
        ins = model.get_input_types() if isinstance(model, InferenceModel) else []
        outs = model.get_output_types() if isinstance(model, InferenceModel) else []
        return {"name": model_name, "platform": "", "inputs": ins, "outputs": outs}

    async def model_ready(self, model_name: str, disable_predictor_health_check: bool = False) -> bool:
        if self._model_registry.get_model(model_name) is None:
            raise ModelNotFound(model_name)
        return await self._model_registry.is_model_ready(model_name)

    def decode(self, body, headers: Optional[Dict[str, str]], protocol_version: str = "v1",
               model_name: Optional[str] = None) -> Tuple[Union[Dict, InferRequest], Dict]:
        """dataplane.py:332-405 — V1: JSON dict. V2: JSON or JSON+binary (Inference-Header-Content-Length)."""
        headers = {k.lower(): v for k, v in (headers or {}).items()}
        if isinstance(body, InferRequest):
            return body, {}
        if protocol_version == "v2":
            if isinstance(body, dict):
                return InferRequest.from_dict(body, model_name), {}
            if "inference-header-content-length" in headers:
                return InferRequest.from_bytes(body, int(headers["inference-header-content-length"]), model_name), {}
            return InferRequest.from_bytes(body, len(body), model_name), {}
        if isinstance(body, (bytes, bytearray)):
            try:
                body = json.loads(body)
            except json.JSONDecodeError as e:
                raise InvalidInput(f"Unrecognized request format: {e}")
        return body, {}

    def encode(self, model_name, response, headers, req_attributes: Dict) -> Tuple[Union[Dict, bytes], Dict[str, str]]:
        """dataplane.py:407-437"""
        response_headers: Dict[str, str] = {}
        if isinstance(response, InferResponse):
            response, json_length = response.to_rest()
            if json_length is not None:
                response_headers["inference-header-content-length"] = str(json_length)
                response_headers["content-type"] = "application/octet-stream"
        return response, response_headers

    async def explain(self, model_name: str, request: Union[Dict, InferRequest],
                      headers: Optional[Dict[str, str]] = None):
        """dataplane.py:477-506: same call as infer() with the EXPLAIN verb"""
        from ..model import InferenceVerb
        response_headers: Dict[str, str] = {}
        model = await self.get_model(model_name)
        if not isinstance(model, InferenceModel):
            raise ValueError(f"Model of type {type(model).__name__} does not support inference")
        response, res_headers = await model(request, headers=headers, verb=InferenceVerb.EXPLAIN)
        response_headers.update(res_headers)
        return response, response_headers

    async def infer(self, model_name: str, request: Union[Dict, InferRequest],
                    headers: Optional[Dict[str, str]] = None):
        """dataplane.py:439-475"""
        response_headers: Dict[str, str] = {}
        model = await self.get_model(model_name)
        if not isinstance(model, InferenceModel):
            raise ValueError(f"Model of type {type(model).__name__} does not support inference")
        response, res_headers = await model(request, headers=headers)
        response_headers.update(res_headers)
        return response, response_headers
