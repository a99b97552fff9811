"""`inference.GRPCInferenceService` over grpc.aio (servicer.py:37-127, server.py of the reference): the handlers call the
same DataPlane as the REST routes; kserve errors map to gRPC status codes (InvalidInput -> INVALID_ARGUMENT,
ModelNotFound -> NOT_FOUND, ModelNotReady -> UNAVAILABLE, NotImplementedError -> UNIMPLEMENTED, else INTERNAL)."""
from typing import Optional

import grpc

from ... import errors as E
from ..dataplane import DataPlane
from ..infer_type import InferResponse
from . import pb
from .convert import infer_request_from_grpc, infer_response_to_grpc

_STATUS = [(E.InvalidInput, grpc.StatusCode.INVALID_ARGUMENT), (E.ModelNotFound, grpc.StatusCode.NOT_FOUND),
           (E.ModelNotReady, grpc.StatusCode.UNAVAILABLE), (NotImplementedError, grpc.StatusCode.UNIMPLEMENTED)]


class InferenceServicer:
    def __init__(self, data_plane: DataPlane):
        self._data_plane = data_plane

    async def ServerLive(self, request, context):
        return pb.ServerLiveResponse(live=(await self._data_plane.live())["status"] == "alive")

    async def ServerReady(self, request, context):
        return pb.ServerReadyResponse(ready=await self._data_plane.ready())

    async def ModelReady(self, request, context):
        return pb.ModelReadyResponse(ready=await self._data_plane.model_ready(model_name=request.name))

    async def ServerMetadata(self, request, context):
        md = self._data_plane.metadata()
        return pb.ServerMetadataResponse(name=md["name"], version=md["version"], extensions=md["extensions"])

    async def ModelMetadata(self, request, context):
        md = await self._data_plane.model_metadata(model_name=request.name)
        return pb.ModelMetadataResponse(name=md["name"], platform=md["platform"], inputs=md["inputs"], outputs=md["outputs"])

    async def RepositoryModelLoad(self, request, context):
        ok = self._data_plane.model_registry.load(request.model_name)
        return pb.RepositoryModelLoadResponse(model_name=request.model_name, isLoaded=bool(ok))

    async def RepositoryModelUnload(self, request, context):
        self._data_plane.model_registry.unload(request.model_name)
        return pb.RepositoryModelUnloadResponse(model_name=request.model_name, isUnloaded=True)

    async def ModelInfer(self, request, context):
        headers = {k: v for k, v in (context.invocation_metadata() or [])} if context is not None else {}
        infer_request = infer_request_from_grpc(request)
        if not await self._data_plane.model_ready(request.model_name, True):
            raise E.ModelNotReady(request.model_name)
        response, _ = await self._data_plane.infer(request=infer_request, headers=headers, model_name=request.model_name)
        if isinstance(response, pb.ModelInferResponse):
            return response
        if isinstance(response, InferResponse):
            if infer_request.use_raw:          # raw tensors in -> raw tensors out (the convention of OIP clients)
                response._use_binary_outputs = True
            return infer_response_to_grpc(response)
        return pb.ModelInferResponse(id=response["id"], model_name=response["model_name"], outputs=response["outputs"])


def _wrap(fn):
    async def handler(request, context):
        try:
            return await fn(request, context)
        except Exception as e:      # noqa: BLE001 - every error becomes a status code, never a dropped connection
            code = next((c for t, c in _STATUS if isinstance(e, t)), grpc.StatusCode.INTERNAL)
            await context.abort(code, getattr(e, "reason", None) or str(e))
    return handler


def generic_handler(servicer: InferenceServicer) -> grpc.GenericRpcHandler:
    table = {}
    for m in pb.METHODS:
        req, res = getattr(pb, m + "Request"), getattr(pb, m + "Response")
        table[m] = grpc.unary_unary_rpc_method_handler(_wrap(getattr(servicer, m)), request_deserializer=req.FromString,
                                                       response_serializer=res.SerializeToString)
    return grpc.method_handlers_generic_handler(pb.SERVICE_NAME, table)


class GRPCServer:
    def __init__(self, port: int, data_plane: DataPlane, host: str = "[::]", max_message_bytes: int = 256 << 20):
        self._port, self._host, self._dp, self._max = port, host, data_plane, max_message_bytes
        self._server: Optional[grpc.aio.Server] = None
        self.bound_port: Optional[int] = None

    async def start(self):
        self._server = grpc.aio.server(options=[("grpc.max_send_message_length", self._max),
                                                ("grpc.max_receive_message_length", self._max)])
        self._server.add_generic_rpc_handlers((generic_handler(InferenceServicer(self._dp)),))
        self.bound_port = self._server.add_insecure_port(f"{self._host}:{self._port}")
        await self._server.start()
        return self

    async def wait(self):
        await self._server.wait_for_termination()

    async def stop(self, grace: float = 0.5):
        if self._server is not None:
            await self._server.stop(grace)
            self._server = None
