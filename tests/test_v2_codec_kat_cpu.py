"""Known-answer tests for the V2 (Open Inference Protocol) REST codec — SURVEY.md §8a row 23.

The byte strings and scenarios are the ones the reference's own test-suite pins
(python/kserve/test/test_infer_type.py:208-545 request side, :723-1090 response side): compact JSON header with keys
in the order id, model_name, inputs[name, shape, datatype, parameters, data], outputs, followed by the raw
little-endian tensors in input order; BYTES elements carry a uint32 length prefix; FP16 must travel as binary data."""
import copy
import json

import numpy as np
import pytest

from kserve_b200.kserve_api.errors import InvalidInput
from kserve_b200.kserve_api.protocol.infer_type import (InferInput, InferOutput, InferRequest, InferResponse,
                                                        RequestedOutput, _contains_fp16_datatype, serialize_byte_tensor)

RID = "4be4e82f-5500-420a-a5c5-ac86841e271b"


def _mixed_request():
    a = InferInput("input1", [3], "INT32", data=np.array([1, 2, 3], dtype=np.int32), parameters={"test-str": "dummy"})
    b = InferInput("input2", [1], "BYTES", parameters={"test-int": 2})
    b.set_data_from_numpy(np.array(["test"], dtype=np.object_), binary_data=True)
    c = InferInput("input3", [3], "FP16")
    c.set_data_from_numpy(np.array([1.2, 2.2, 3.2], dtype=np.float16), binary_data=True)
    outs = [RequestedOutput(n, {"test-str": "dummy", "test-bool": True, "test-int": 100}) for n in ("output-0", "output-1")]
    return InferRequest(request_id=RID, model_name="test_model", infer_inputs=[a, b, c], request_outputs=outs)


def test_request_to_rest_mixed_binary_known_answer():
    body, json_length = _mixed_request().to_rest()        # test_infer_type.py:208-268
    assert body == (b'{"id":"4be4e82f-5500-420a-a5c5-ac86841e271b","model_name":"test_model","inputs":[{"name":"input1","shape":[3],'
                    b'"datatype":"INT32","parameters":{"test-str":"dummy"},"data":[1,2,3]},{"name":"input2","shape":[1],'
                    b'"datatype":"BYTES","parameters":{"test-int":2,"binary_data_size":8}},{"name":"input3","shape":[3],'
                    b'"datatype":"FP16","parameters":{"binary_data_size":6}}],"outputs":[{"name":"output-0","parameters":'
                    b'{"test-str":"dummy","test-bool":true,"test-int":100}},{"name":"output-1","parameters":{"test-str":"dummy",'
                    b'"test-bool":true,"test-int":100}}]}\x04\x00\x00\x00test\xcd<f@fB')
    assert json_length == 546


def test_request_without_binary_data_is_a_dict_and_fp16_needs_binary():
    r = InferRequest(request_id=RID, model_name="test_model", infer_inputs=[
        InferInput("input1", [3], "INT32", data=[1, 2, 3]), InferInput("input2", [1], "BYTES", data=["test"]),
        InferInput("input3", [3], "FP32", data=[1.2, 2.2, 3.2], parameters={"test-int": 2})])
    body, n = r.to_rest()                                  # :306-391
    assert n is None and body["inputs"][2] == {"name": "input3", "shape": [3], "datatype": "FP32", "parameters": {"test-int": 2},
                                               "data": [1.2, 2.2, 3.2]}
    bad = InferInput("x", [3], "FP16")
    bad.set_data_from_numpy(np.array([1.2, 2.2, 3.2], dtype=np.float16), binary_data=False)   # accepted here ...
    with pytest.raises(InvalidInput):                      # ... rejected when the REST body is built (:270-304)
        InferRequest(model_name="m", infer_inputs=[bad]).to_rest()
    with pytest.raises(InvalidInput):                      # :416-428
        InferRequest(model_name="m", infer_inputs=[InferInput("y", [1], "INT32")]).to_rest()


def test_request_from_bytes_round_trip_and_errors():
    r = _mixed_request()
    expected = copy.deepcopy(r)
    body, n = r.to_rest()
    back = InferRequest.from_bytes(body, n, "test_model")  # :430-510
    for i in (1, 2):
        back.inputs[i].set_data_from_numpy(back.inputs[i].as_numpy(), binary_data=True)
    back.inputs[0].data = np.array(back.inputs[0].data, dtype=np.int32)
    assert back == expected and back.request_outputs == expected.request_outputs
    assert back.inputs[2].as_numpy().dtype == np.float16
    trunc = b'{"id": "1", "inputs": [{"name": "input1", "shape": [1], "datatype": "INT32", "data": [1]}'
    with pytest.raises(InvalidInput):                      # :512-518 invalid JSON
        InferRequest.from_bytes(trunc, 100, "test_model")
    hdr = b'{"id":"509c5da9-80d4-46e8-a50c-0bba2b9d76f8","inputs":[{"name":"input1","shape":[3],"datatype":"INT32"}]}'
    with pytest.raises(InvalidInput):                      # :527-537 raw tensor without binary_data_size
        InferRequest.from_bytes(hdr + b"\x01\x00\x00\x00\x02\x00\x00\x00\x03\x00\x00\x00", len(hdr), "test_model")
    fp16_json = b'{"id": "1", "inputs": [{"name": "input1", "shape": [1], "datatype": "FP16", "data": [1]}]}'
    with pytest.raises(InvalidInput):                      # :539-544 FP16 via JSON
        InferRequest.from_bytes(fp16_json, len(fp16_json), "test_model")


def test_response_from_bytes_and_binary_bytes_tensor():
    plain = b'{"id": "1", "model_name": "test_model", "outputs": [{"name": "output1", "shape": [1], "datatype": "INT32", "data": [1]}]}'
    r = InferResponse.from_bytes(plain, len(plain))        # :998-1014
    assert (r.id, r.model_name, r.outputs[0].name, r.outputs[0].shape, r.outputs[0].datatype, r.outputs[0].data) == \
        ("1", "test_model", "output1", [1], "INT32", [1])
    raw = serialize_byte_tensor(np.array([b"cat", b"dog", b"bird", b"fish"], dtype=np.object_)).item()
    assert raw == b"\x03\x00\x00\x00cat\x03\x00\x00\x00dog\x04\x00\x00\x00bird\x04\x00\x00\x00fish"
    hdr = json.dumps({"model_name": "test_model", "id": "1", "outputs": [
        {"name": "output1", "shape": [4], "datatype": "BYTES", "parameters": {"binary_data_size": len(raw)}}]}).encode()
    r = InferResponse.from_bytes(hdr + raw, len(hdr))      # :1016-1048
    assert r.outputs[0].data == ["cat", "dog", "bird", "fish"]
    missing = b'{"id": "1", "model_name": "test_model", "outputs": [{"name": "output1", "shape": [1], "datatype": "INT32"}]}'
    with pytest.raises(InvalidInput):                      # :1050-1055
        InferResponse.from_bytes(missing, len(missing))


def test_response_to_rest_binary_selection_and_helpers():
    o1 = InferOutput("output1", [2], "INT32", data=[1, 2])
    o2 = InferOutput("output2", [2], "FP32", data=[0.5, 1.5])
    # request-level binary_data_output=True -> every output binary (:867-904); per-output `binary_data` wins (:906-950)
    r = InferResponse("1", "m", [o1, o2], use_binary_outputs=True,
                      requested_outputs=[RequestedOutput("output1", {"binary_data": False}), RequestedOutput("output2")])
    body, n = r.to_rest()
    head = json.loads(body[:n])
    assert head["outputs"][0]["data"] == [1, 2] and "data" not in head["outputs"][1]
    assert head["outputs"][1]["parameters"]["binary_data_size"] == 8
    assert np.frombuffer(body[n:], dtype=np.float32).tolist() == [0.5, 1.5]
    back = InferResponse.from_bytes(body, n)
    assert back.outputs[1].data.tolist() == [0.5, 1.5]
    assert r.get_output_by_name("output2") == o2 and r.get_output_by_name("nope") is None     # :1057-1092
    assert _contains_fp16_datatype(InferResponse("1", "m", [InferOutput("h", [1], "FP16", data=[1]), o1])) is True
    assert _contains_fp16_datatype(InferResponse("1", "m", [o1])) is False
    assert _contains_fp16_datatype(InferResponse("1", "m", [])) is False                       # :1094-1120
