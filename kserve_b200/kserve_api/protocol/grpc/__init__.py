"""gRPC front-end of the Open Inference Protocol (`inference.GRPCInferenceService`) in front of the same
`Model.predict()` boundary as the REST routes — SURVEY.md §8(f) rank 2
(python/kserve/kserve/protocol/grpc/{grpc_predict_v2.proto, servicer.py:37-127, server.py})."""
from . import pb  # noqa: F401
from .server import GRPCServer, InferenceServicer  # noqa: F401
