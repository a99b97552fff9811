"""REST-level known answers for the V1 / V2 routes, restating the scenarios and the exact
response bytes that the reference's own suite pins (python/kserve/test/test_server.py:415-485 V1, :487-870 V2):
status codes and error bodies, compact JSON, the JSON-only response shape (model_name, model_version, id, parameters,
outputs[name, shape, datatype, parameters, data]) versus the binary-extension shape (id, model_name, model_version,
outputs + raw tensors + Inference-Header-Content-Length), FP16 rules and binary_data precedence."""
import json

import numpy as np
import pytest
from fastapi.testclient import TestClient

from kserve_b200.kserve_api import Model, ModelServer
from kserve_b200.kserve_api.protocol.infer_type import (InferInput, InferRequest, RequestedOutput, get_predict_input, get_predict_response)

HDR = "inference-header-content-length"


class Echo(Model):
    """test_server.py:146-176"""
    def __init__(self, name):
        super().__init__(name)
        self.ready = True

    async def predict(self, request, headers=None):
        if isinstance(request, InferRequest):
            res = get_predict_response(request, get_predict_input(request), self.name)
            if request.parameters:
                res.parameters = request.parameters
            if request.inputs[0].parameters:
                res.outputs[0].parameters = request.inputs[0].parameters
            return res
        return {"predictions": request["inputs"] if "inputs" in request else request["instances"]}

    async def explain(self, request, headers=None):
        return {"predictions": request["inputs"] if "inputs" in request else request["instances"]}


class FP16Out(Model):
    """test_server.py:273-303: one FP16 and one FP32 output from an FP32 input"""
    def __init__(self, name):
        super().__init__(name)
        self.ready = True

    async def predict(self, request, headers=None):
        x = request.get_input_by_name("fp32_input").as_numpy()
        res = get_predict_response(request, {"fp16_output": x.astype(np.float16).flatten(), "fp32_output": x.flatten()}, self.name)
        if request.parameters:
            res.parameters = request.parameters
            res.parameters.pop("binary_data_output", None)
        return res


class FP16In(Model):
    """test_server.py:306-336"""
    def __init__(self, name):
        super().__init__(name)
        self.ready = True

    async def predict(self, request, headers=None):
        res = get_predict_response(request, {
            "str_output": request.get_input_by_name("str_input").as_numpy().flatten(),
            "fp32_output": request.get_input_by_name("fp16_input").as_numpy().astype(np.float32).flatten()}, self.name)
        return res


@pytest.fixture(scope="module")
def client():
    app = ModelServer().create_application([Echo("TestModel"), FP16In("FP16InputModel"), FP16Out("FP16OutputModel")])
    with TestClient(app) as c:
        yield c


FP32 = np.array([[6.8, 2.8, 4.8, 1.4], [6.0, 3.4, 4.5, 1.6]], dtype=np.float32)
FP32_JSON = b"[6.800000190734863,2.799999952316284,4.800000190734863,1.399999976158142,6.0,3.4000000953674316,4.5,1.600000023841858]"
FP16_RAW = b"\xcdF\x9aA\xcdD\x9a=\x00F\xcdB\x80Df>"
FP32_RAW = b"\x9a\x99\xd9@333@\x9a\x99\x99@33\xb3?\x00\x00\xc0@\x9a\x99Y@\x00\x00\x90@\xcd\xcc\xcc?"


def _fp16out_request(**kw):
    return InferRequest(model_name="FP16OutputModel", request_id="123",
                        infer_inputs=[InferInput("fp32_input", [2, 4], "FP32", data=FP32.tolist())], **kw).to_rest()[0]


def test_v1_routes(client):
    assert client.get("/").json() == {"status": "alive"}
    assert client.get("/v1/models/TestModel").status_code == 200
    r = client.get("/v1/models/InvalidModel")
    assert r.status_code == 404 and r.json() == {"error": "Model with name InvalidModel does not exist."}
    assert client.get("/v1/models").json() == {"models": ["TestModel", "FP16InputModel", "FP16OutputModel"]}
    for verb in ("predict", "explain"):
        r = client.post(f"/v1/models/TestModel:{verb}", content=b'{"instances":[[1,2]]}')
        assert r.status_code == 200 and r.content == b'{"predictions":[[1,2]]}' and r.headers["content-type"] == "application/json"
    r = client.get("/unknown_path")
    assert r.status_code == 404 and r.json() == {"detail": "Not Found"}
    assert client.get("/metrics").status_code == 200


def test_v2_infer_json_and_parameters(client):
    assert client.get("/v2/models").json() == {"models": ["TestModel", "FP16InputModel", "FP16OutputModel"]}
    body = b'{"inputs": [{"name": "input-0","shape": [1, 2],"datatype": "INT32","data": [[1,2]]}]}'
    r = client.post("/v2/models/TestModel/infer", content=body, headers={"content-type": "application/json"})
    assert r.status_code == 200 and r.headers["content-type"] == "application/json"
    assert json.loads(r.content)["outputs"][0]["data"] == [1, 2]
    p = {"test-str": "dummy", "test-bool": True, "test-int": 100, "test-float": 1.3}
    req = InferRequest(model_name="TestModel", request_id="123", parameters=dict(p),
                       infer_inputs=[InferInput("input-0", [1, 2], "INT32", data=[1, 2], parameters=dict(p))])
    r = client.post("/v2/models/TestModel/infer", content=json.dumps(req.to_rest()[0]).encode())
    assert r.status_code == 200
    got = json.loads(r.content)
    assert (got["id"], got["model_name"], got["parameters"]) == ("123", "TestModel", p)
    assert got["outputs"] == [{"name": "output-0", "shape": [1, 2], "datatype": "INT32", "parameters": p, "data": [1, 2]}]


def test_v2_fp16_input(client):
    fp16 = FP32.astype(np.float16)
    strs = [["cat", "dog", "cat", "dog"], ["cat", "dog", "cat", "dog"]]
    a = InferInput("fp16_input", [2, 4], "FP16")
    a.set_data_from_numpy(fp16, binary_data=True)
    body, n = InferRequest(model_name="FP16InputModel", request_id="123",
                           infer_inputs=[a, InferInput("str_input", [2, 4], "BYTES", data=strs)]).to_rest()
    r = client.post("/v2/models/FP16InputModel/infer", content=body,
                    headers={HDR: str(n), "Content-Type": "application/octet-stream"})
    assert r.status_code == 200                       # :591-632 JSON-only response shape
    assert r.content == (b'{"model_name":"FP16InputModel","model_version":null,"id":"123","parameters":null,"outputs":['
                         b'{"name":"str_output","shape":[8],"datatype":"BYTES","parameters":null,"data":["cat","dog","cat","dog","cat","dog","cat","dog"]},'
                         b'{"name":"fp32_output","shape":[8],"datatype":"FP32","parameters":null,"data":[6.80078125,2.80078125,4.80078125,1.400390625,6.0,3.400390625,4.5,1.599609375]}]}')
    bad = {"model_name": "FP16InputModel", "request_id": "123", "inputs": [
        {"name": "fp16_input", "shape": [2, 4], "datatype": "FP16", "data": fp16.tolist()},
        {"name": "str_input", "shape": [2, 4], "datatype": "BYTES", "data": strs}]}
    assert client.post("/v2/models/FP16InputModel/infer", json=bad).status_code == 400      # :634-667 FP16 via JSON


def test_v2_fp16_output_and_binary_precedence(client):
    mixed = (b'{"id":"123","model_name":"FP16OutputModel","model_version":null,"outputs":[{"name":"fp16_output","shape":[8],'
             b'"datatype":"FP16","parameters":{"binary_data_size":16}},{"name":"fp32_output","shape":[8],"datatype":"FP32","data":'
             + FP32_JSON + b'}]}' + FP16_RAW)
    outs = [RequestedOutput("fp16_output", {"binary_data": True}), RequestedOutput("fp32_output", {"binary_data": False})]
    r = client.post("/v2/models/FP16OutputModel/infer", json=_fp16out_request(request_outputs=outs))
    assert r.status_code == 200 and r.content == mixed and r.headers.get(HDR) == "345"       # :669-708
    # FP16 output requested without the binary format -> 400 (:710-742)
    bad = {"model_name": "FP16OutputModel", "request_id": "123",
           "inputs": [{"name": "fp32_input", "shape": [2, 4], "datatype": "FP32", "data": FP32.tolist()}],
           "outputs": [{"name": "fp16_output"}, {"name": "fp32_output", "parameters": {"binary_data": False}}]}
    assert client.post("/v2/models/FP16OutputModel/infer", json=bad).status_code == 400
    # only the requested output comes back (:744-778)
    r = client.post("/v2/models/FP16OutputModel/infer",
                    json=_fp16out_request(request_outputs=[RequestedOutput("fp32_output", {"binary_data": False})]))
    assert r.status_code == 200
    assert r.content == (b'{"model_name":"FP16OutputModel","model_version":null,"id":"123","parameters":null,"outputs":['
                         b'{"name":"fp32_output","shape":[8],"datatype":"FP32","parameters":null,"data":' + FP32_JSON + b'}]}')
    # request-level binary_data_output (:780-810) ...
    r = client.post("/v2/models/FP16OutputModel/infer", json=_fp16out_request(parameters={"binary_data_output": True}))
    assert r.status_code == 200 and r.headers.get(HDR) == "256"
    assert r.content == (b'{"id":"123","model_name":"FP16OutputModel","model_version":null,"outputs":[{"name":"fp16_output","shape":[8],'
                         b'"datatype":"FP16","parameters":{"binary_data_size":16}},{"name":"fp32_output","shape":[8],"datatype":"FP32",'
                         b'"parameters":{"binary_data_size":32}}]}' + FP16_RAW + FP32_RAW)
    # ... which a per-output binary_data overrides (:812-852)
    r = client.post("/v2/models/FP16OutputModel/infer",
                    json=_fp16out_request(parameters={"binary_data_output": True}, request_outputs=outs))
    assert r.status_code == 200 and r.content == mixed and r.headers.get(HDR) == "345"
