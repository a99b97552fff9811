"""kserve.Model public interface (mirrors python/kserve/kserve/model.py:44-469): a model server calls
``await model(body, headers)`` which runs preprocess -> validate -> predict -> postprocess, each timed."""
from __future__ import annotations

import inspect
import logging
import time
from abc import ABC, abstractmethod
from enum import Enum
from typing import Dict, List, Optional

from .errors import InvalidInput
from .metrics import EXPLAIN_HIST_TIME, POST_HIST_TIME, PRE_HIST_TIME, PREDICT_HIST_TIME, get_labels
from .protocol.infer_type import InferRequest

trace_logger = logging.getLogger("kserve.trace")


class BaseKServeModel(ABC):
    """model.py:44-96 — name / ready / engine + lifecycle hooks the model server drives."""

    @abstractmethod
    def __init__(self, name: str):
        self.name = name
        self.ready = False
        self.engine = False

    async def healthy(self) -> bool:
        return self.ready

    def load(self) -> bool:
        self.ready = True
        return self.ready

    def start(self):
        self.ready = True

    async def start_engine(self):
        self.ready = True

    def stop(self):
        self.ready = False

    def stop_engine(self):
        self.ready = False


class InferenceVerb(Enum):
    EXPLAIN = 1
    PREDICT = 2


class InferenceModel(BaseKServeModel):
    """model.py:108-136"""

    @abstractmethod
    def __call__(self, body, headers: Optional[Dict[str, str]] = None, verb: InferenceVerb = InferenceVerb.PREDICT):
        pass

    def get_input_types(self) -> List[Dict]:
        return []

    def get_output_types(self) -> List[Dict]:
        return []


def get_latency_ms(start: float, end: float) -> float:
    return round((end - start) * 1000, 9)


async def _maybe_await(fn, *args):
    r = fn(*args)
    if inspect.isawaitable(r):
        r = await r
    return r


class Model(InferenceModel):
    def __init__(self, name: str, return_response_headers: bool = False):
        super().__init__(name)
        self.enable_latency_logging = False
        self.required_response_headers = return_response_headers

    async def __call__(self, body, headers: Optional[Dict[str, str]] = None,
                       verb: InferenceVerb = InferenceVerb.PREDICT):
        """model.py:173-259: returns (response, response_headers)."""
        request_id = headers.get("x-request-id", "N.A.") if headers else "N.A."
        preprocess_ms = explain_ms = predict_ms = postprocess_ms = 0
        labels = get_labels(self.name)
        response_headers: Dict[str, str] = {}
        with PRE_HIST_TIME.labels(**labels).time():
            start = time.time()
            payload = await _maybe_await(self.preprocess, body, headers)
            preprocess_ms = get_latency_ms(start, time.time())
        payload = self.validate(payload)
        extra = (response_headers,) if self.required_response_headers else ()
        if verb == InferenceVerb.EXPLAIN:
            with EXPLAIN_HIST_TIME.labels(**labels).time():
                start = time.time()
                response = await _maybe_await(self.explain, payload, headers)
                explain_ms = get_latency_ms(start, time.time())
        elif verb == InferenceVerb.PREDICT:
            with PREDICT_HIST_TIME.labels(**labels).time():
                start = time.time()
                response = await _maybe_await(self.predict, payload, headers, *extra)
                predict_ms = get_latency_ms(start, time.time())
        else:
            raise NotImplementedError
        with POST_HIST_TIME.labels(**labels).time():
            start = time.time()
            response = await _maybe_await(self.postprocess, response, headers, *extra)
            postprocess_ms = get_latency_ms(start, time.time())
        if self.enable_latency_logging is True:
            trace_logger.info(f"requestId: {request_id}, preprocess_ms: {preprocess_ms}, explain_ms: {explain_ms}, "
                              f"predict_ms: {predict_ms}, postprocess_ms: {postprocess_ms}")
        return response, response_headers

    def validate(self, payload):
        """model.py:298-324 (protocol-specific shape checks)."""
        if isinstance(payload, InferRequest):
            return payload
        if isinstance(payload, dict):
            if "instances" in payload and not isinstance(payload["instances"], list):
                raise InvalidInput('Expected "instances" to be a list')
            if "inputs" in payload and not isinstance(payload["inputs"], list):
                raise InvalidInput('Expected "inputs" to be a list')
        return payload

    async def preprocess(self, payload, headers: Dict[str, str] = None):
        return payload

    async def postprocess(self, result, headers: Dict[str, str] = None, response_headers: Dict[str, str] = None):
        return result

    async def predict(self, payload, headers: Dict[str, str] = None, response_headers: Dict[str, str] = None):
        """model.py:403-430: the reference default forwards to a remote predictor; a runtime model overrides it."""
        raise NotImplementedError("predict() is not implemented by this model")

    async def explain(self, payload, headers: Dict[str, str] = None):
        raise NotImplementedError("explain() is not implemented by this model")
