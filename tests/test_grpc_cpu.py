"""Open Inference Protocol over gRPC against an in-process grpc.aio server, restating the
scenarios of the reference's python/kserve/test/test_grpc_server.py:158-660: typed `contents` in -> typed `contents`
out, raw tensors in -> raw tensors out, FP16 only as raw bytes, the two validation errors of servicer.py:37-50, plus the
health / metadata RPCs of servicer.py:52-88."""
import asyncio

import grpc
import numpy as np
import pytest
from google.protobuf.json_format import MessageToDict

from kserve_b200.kserve_api import Model
from kserve_b200.kserve_api.model_repository import ModelRepository
from kserve_b200.kserve_api.protocol.grpc import GRPCServer, pb
from kserve_b200.kserve_api.protocol.infer_type import InferRequest, get_predict_response
from kserve_b200.kserve_api.protocol.rest.openai.dataplane import OpenAIDataPlane


class Renamer(Model):
    """every input tensor comes back flattened under <prefix>_output (test_grpc_server.py:38-66)"""
    def __init__(self, name):
        super().__init__(name)
        self.ready = True

    async def predict(self, request: InferRequest, headers=None):
        outs = {i.name.replace("_input", "_output"): i.as_numpy().flatten() for i in request.inputs}
        res = get_predict_response(request, outs, self.name)
        if request.parameters:
            res.parameters = request.parameters
        return res


class Halver(Model):
    """FP32 in -> FP16 + FP32 out (test_grpc_server.py:69-97)"""
    def __init__(self, name):
        super().__init__(name)
        self.ready = True

    async def predict(self, request: InferRequest, headers=None):
        x = request.get_input_by_name("fp32_input").as_numpy()
        return get_predict_response(request, {"fp16_output": x.astype(np.float16).flatten(), "fp32_output": x.flatten()}, self.name)


def _call(coro_fn):
    async def main():
        repo = ModelRepository()
        repo.update(Renamer("TestModel"))
        repo.update(Halver("FP16OutputModel"))
        srv = await GRPCServer(0, OpenAIDataPlane(model_registry=repo), host="127.0.0.1").start()
        try:
            async with grpc.aio.insecure_channel(f"127.0.0.1:{srv.bound_port}") as ch:
                def rpc(name):
                    return ch.unary_unary(f"/{pb.SERVICE_NAME}/{name}", request_serializer=getattr(pb, name + "Request").SerializeToString,
                                          response_deserializer=getattr(pb, name + "Response").FromString)
                return await coro_fn(rpc)
        finally:
            await srv.stop(0)
    return asyncio.run(main())


FP32 = [6.8, 2.8, 4.8, 1.4, 6.0, 3.4, 4.5, 1.6]


def test_typed_contents_round_trip():
    req = pb.ModelInferRequest(model_name="TestModel", id="123", inputs=[
        {"name": "fp32_input", "shape": [2, 4], "datatype": "FP32", "contents": {"fp32_contents": FP32}},
        {"name": "int32_input", "shape": [2, 4], "datatype": "INT32", "contents": {"int_contents": [6, 2, 4, 1, 6, 3, 4, 1]}},
        {"name": "string_input", "shape": [8], "datatype": "BYTES",
         "contents": {"bytes_contents": [b"Cat", b"Dog", b"Wolf", b"Cat", b"Dog", b"Wolf", b"Dog", b"Wolf"]}},
        {"name": "uint8_input", "shape": [2, 4], "datatype": "UINT8", "contents": {"uint_contents": [6, 2, 4, 1, 6, 3, 4, 1]}},
        {"name": "bool_input", "shape": [8], "datatype": "BOOL", "contents": {"bool_contents": [True, False] * 4}}])
    res = _call(lambda rpc: rpc("ModelInfer")(req))
    d = MessageToDict(res, preserving_proto_field_name=True)          # test_grpc_server.py:233-297
    assert d["model_name"] == "TestModel" and d["id"] == "123"
    by = {o["name"]: o for o in d["outputs"]}
    assert by["fp32_output"] == {"name": "fp32_output", "datatype": "FP32", "shape": ["8"],
                                 "contents": {"fp32_contents": pytest.approx(FP32, rel=1e-6)}}
    assert by["int32_output"]["contents"] == {"int_contents": [6, 2, 4, 1, 6, 3, 4, 1]} and by["int32_output"]["datatype"] == "INT32"
    assert by["string_output"]["contents"] == {"bytes_contents": ["Q2F0", "RG9n", "V29sZg==", "Q2F0", "RG9n", "V29sZg==", "RG9n", "V29sZg=="]}
    assert by["uint8_output"]["contents"] == {"uint_contents": [6, 2, 4, 1, 6, 3, 4, 1]} and by["uint8_output"]["datatype"] == "UINT8"
    assert by["bool_output"]["contents"] == {"bool_contents": [True, False] * 4}


def test_raw_tensors_in_raw_tensors_out_and_parameters():
    ids = np.arange(12, dtype=np.int64).reshape(3, 4)
    req = pb.ModelInferRequest(model_name="TestModel", id="7", parameters={"max_tokens": {"int64_param": 5}, "tag": {"string_param": "x"}},
                               inputs=[{"name": "ids_input", "shape": [3, 4], "datatype": "INT64"}], raw_input_contents=[ids.tobytes()])
    res = _call(lambda rpc: rpc("ModelInfer")(req))
    assert len(res.raw_output_contents) == 1 and not res.outputs[0].HasField("contents")
    assert np.array_equal(np.frombuffer(res.raw_output_contents[0], dtype=np.int64), ids.reshape(-1))
    assert list(res.outputs[0].shape) == [12] and res.outputs[0].datatype == "INT64"
    assert res.parameters["max_tokens"].int64_param == 5 and res.parameters["tag"].string_param == "x"


def test_fp16_travels_as_raw_bytes():
    req = pb.ModelInferRequest(model_name="FP16OutputModel", id="123", inputs=[
        {"name": "fp32_input", "shape": [2, 4], "datatype": "FP32", "contents": {"fp32_contents": FP32}}])
    res = _call(lambda rpc: rpc("ModelInfer")(req))                  # test_grpc_server.py:427-495
    assert [o.name for o in res.outputs] == ["fp16_output", "fp32_output"] and len(res.raw_output_contents) == 2
    assert res.raw_output_contents[0] == b"\xcdF\x9aA\xcdD\x9a=\x00F\xcdB\x80Df>"
    assert np.frombuffer(res.raw_output_contents[1], dtype=np.float32).tolist() == np.array(FP32, dtype=np.float32).tolist()


def test_validation_errors_and_status_codes():
    async def run(rpc):
        out = {}
        bad = [pb.ModelInferRequest(model_name="TestModel", inputs=[{"name": "a_input", "shape": [1], "datatype": "INT32"},
                                                                    {"name": "b_input", "shape": [1], "datatype": "INT32"}],
                                    raw_input_contents=[b"\\x01\\x00\\x00\\x00"]),                         # :562-611
               pb.ModelInferRequest(model_name="TestModel", inputs=[{"name": "a_input", "shape": [1], "datatype": "INT32",
                                                                     "contents": {"int_contents": [1]}}],
                                    raw_input_contents=[b"\\x01\\x00\\x00\\x00"]),                         # :613-660
               pb.ModelInferRequest(model_name="Nope", inputs=[{"name": "a_input", "shape": [1], "datatype": "INT32",
                                                                "contents": {"int_contents": [1]}}])]
        for i, r in enumerate(bad):
            try:
                await rpc("ModelInfer")(r)
                out[i] = None
            except grpc.aio.AioRpcError as e:
                out[i] = (e.code(), e.details())
        out["live"] = (await rpc("ServerLive")(pb.ServerLiveRequest())).live
        out["ready"] = (await rpc("ServerReady")(pb.ServerReadyRequest())).ready
        out["model_ready"] = (await rpc("ModelReady")(pb.ModelReadyRequest(name="TestModel"))).ready
        md = await rpc("ServerMetadata")(pb.ServerMetadataRequest())
        out["md"] = (md.name, list(md.extensions))
        out["mm"] = (await rpc("ModelMetadata")(pb.ModelMetadataRequest(name="TestModel"))).name
        return out
    out = _call(run)
    assert out[0][0] == grpc.StatusCode.INVALID_ARGUMENT and "does not match the expected number of raw input contents (1)" in out[0][1]
    assert out[1][0] == grpc.StatusCode.INVALID_ARGUMENT and "contents field must not be specified when using raw_input_contents" in out[1][1]
    assert out[2][0] == grpc.StatusCode.NOT_FOUND and out[2][1] == "Model with name Nope does not exist."
    assert out["live"] and out["ready"] and out["model_ready"]
    assert out["md"] == ("kserve", ["model_repository_extension"]) and out["mm"] == "TestModel"
