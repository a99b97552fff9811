"""Signature-compatible mirror of the parts of the `kserve` SDK on the LLM predict path
(python/kserve/kserve): the real package cannot be imported in this environment (cloudevents, orjson,
timing_asgi, kubernetes are absent — SURVEY.md §8c), so the plug-in surface is restated here with the same
names, argument meaning and error behaviour."""
from .errors import InferenceError, InvalidInput, ModelNotFound, ModelNotReady  # noqa: F401
from .model import BaseKServeModel, InferenceModel, InferenceVerb, Model  # noqa: F401
from .model_repository import ModelRepository  # noqa: F401
from .model_server import ModelServer  # noqa: F401
from .protocol.infer_type import InferInput, InferOutput, InferRequest, InferResponse  # noqa: F401
