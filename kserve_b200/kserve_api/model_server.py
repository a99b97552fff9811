"""ModelServer (mirrors python/kserve/kserve/model_server.py:48-461): argparse flags, model registration,
uvicorn REST server on --http_port and, with --enable_grpc, the Open Inference Protocol gRPC service on --grpc_port
(protocol/grpc/).  Multi-process workers are outside this runtime's scope (one engine per GPU)."""
from __future__ import annotations

import argparse
import asyncio
import logging
from typing import List, Optional

from .model import BaseKServeModel
from .model_repository import ModelRepository
from .protocol.rest.openai.dataplane import OpenAIDataPlane
from .protocol.rest.server import create_application

DEFAULT_HTTP_PORT = 8080
DEFAULT_GRPC_PORT = 8081

parser = argparse.ArgumentParser(add_help=False)
parser.add_argument("--http_port", default=DEFAULT_HTTP_PORT, type=int, help="The HTTP Port listened to by the model server.")
parser.add_argument("--grpc_port", default=DEFAULT_GRPC_PORT, type=int, help="The gRPC Port listened to by the model server.")
parser.add_argument("--enable_grpc", default=False, type=lambda x: str(x).lower() == "true", help="Enable the gRPC server.")
parser.add_argument("--workers", default=1, type=int, help="Only 1 is supported: one CUDA engine per GPU.")
parser.add_argument("--enable_latency_logging", default=True, type=lambda x: str(x).lower() == "true")
parser.add_argument("--log_config_file", default=None, type=str)
parser.add_argument("--access_log_format", default=None, type=str)
parser.add_argument("--model_name", default="model", type=str, help="The name of the model used on the endpoint path.")
parser.add_argument("--predictor_host", default=None, type=str)
parser.add_argument("--enable_docs_url", default=False, type=lambda x: str(x).lower() == "true")

logger = logging.getLogger("kserve")


class ModelServer:
    def __init__(self, http_port: int = DEFAULT_HTTP_PORT, workers: int = 1,
                 registered_models: Optional[ModelRepository] = None, enable_latency_logging: bool = True,
                 access_log_format: Optional[str] = None, grpc_port: int = DEFAULT_GRPC_PORT, enable_grpc: bool = False,
                 batcher: Optional[tuple] = None):
        if workers != 1:
            raise ValueError("kserve_b200 runs one engine per GPU: --workers must be 1")
        self.http_port = http_port
        self.grpc_port, self.enable_grpc = grpc_port, enable_grpc
        self.batcher = batcher          # (max_batchsize, max_latency_ms): the agent's batcher in front of V1 :predict
        self._grpc_server = None
        self.registered_models = registered_models or ModelRepository()
        self.enable_latency_logging = enable_latency_logging
        self.access_log_format = access_log_format
        self.dataplane = OpenAIDataPlane(model_registry=self.registered_models)
        self._server = None

    def register_model(self, model: BaseKServeModel, name: Optional[str] = None):
        """model_server.py:427-439"""
        if not model.name:
            raise Exception("Failed to register model, model.name must be provided.")
        name = name or model.name
        self.registered_models.update(model, name)
        if hasattr(model, "enable_latency_logging"):
            model.enable_latency_logging = self.enable_latency_logging
        logger.info("Registering model: %s", name)

    def _register_and_check(self, models: List[BaseKServeModel]):
        """model_server.py:441-459: every model must be ready (or be an engine started later)."""
        for model in models:
            if not isinstance(model, BaseKServeModel):
                raise RuntimeError("Model type should be 'BaseKServeModel'")
            if model.ready or model.engine:
                self.register_model(model)
            else:
                raise RuntimeError(f"Failed to start model server, model {model.name} is not ready.")

    def create_application(self, models: List[BaseKServeModel]):
        self._register_and_check(models)
        return create_application(self.dataplane, batcher=self.batcher)

    async def _serve(self, models: List[BaseKServeModel]):
        import uvicorn
        app = self.create_application(models)
        for m in models:
            if m.engine:
                await m.start_engine()
        if self.enable_grpc:
            from .protocol.grpc import GRPCServer
            self._grpc_server = await GRPCServer(self.grpc_port, self.dataplane).start()
        cfg = uvicorn.Config(app, host="0.0.0.0", port=self.http_port, log_level="info", access_log=True)
        self._server = uvicorn.Server(cfg)
        try:
            await self._server.serve()
        finally:
            if self._grpc_server is not None:
                await self._grpc_server.stop()

    def start(self, models: List[BaseKServeModel]):
        """model_server.py:332-377"""
        try:
            asyncio.run(self._serve(models))
        finally:
            self.stop()

    def stop(self, sig: Optional[int] = None):
        for m in list(self.registered_models.get_models().values()):
            try:
                m.stop()
                m.stop_engine()
            except Exception:  # pragma: no cover
                pass
        if self._server is not None:
            self._server.should_exit = True
