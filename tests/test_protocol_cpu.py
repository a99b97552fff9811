"""Host-side protocol plumbing (no GPU): V1/V2 envelopes, binary tensor extension, OpenAI routes, SSE framing,
chat adapter and error mapping, exercised through FastAPI's TestClient with dummy models — the approach of
the reference's python/kserve/test/test_server.py:146-175 and test_openai_completion.py:68-99."""
import json

import numpy as np
import pytest
from fastapi.testclient import TestClient

from kserve_b200.kserve_api import InferInput, InferRequest, Model, ModelServer
from kserve_b200.kserve_api.errors import InvalidInput
from kserve_b200.kserve_api.protocol.infer_type import get_predict_input, get_predict_response
from kserve_b200.kserve_api.protocol.rest.openai.errors import OpenAIError
from kserve_b200.kserve_api.protocol.rest.openai.openai_chat_adapter_model import OpenAIChatAdapterModel
from kserve_b200.kserve_api.protocol.rest.openai.openai_model import ChatPrompt
from kserve_b200.kserve_api.protocol.rest.openai.types import (Completion, CompletionChoice, CompletionChunk,
                                                               CompletionChunkChoice, CompletionRequest, UsageInfo)


class EchoModel(Model):
    """test_server.py:146-175 DummyModel: echoes instances / inputs."""

    def __init__(self, name):
        super().__init__(name)
        self.ready = True

    async def predict(self, payload, headers=None, response_headers=None):
        if isinstance(payload, InferRequest):
            arr = get_predict_input(payload)
            return get_predict_response(payload, arr, self.name)
        return {"predictions": payload["instances"]}


class DummyOpenAI(OpenAIChatAdapterModel):
    """test_openai_completion.py:68-99: replays a canned completion; streams it in pieces."""

    def __init__(self, name):
        super().__init__(name)

    def apply_chat_template(self, request):
        return ChatPrompt(prompt="".join(f"<|{m.role}|>{m.content}\n" for m in request.messages))

    async def create_completion(self, request: CompletionRequest, raw_request=None, context=None):
        if request.prompt == "boom":
            raise OpenAIError("'frequency_penalty' is not supported")
        text = f"echo:{request.prompt}"
        if request.stream:
            async def gen():
                for piece in (text[:3], text[3:]):
                    yield "data: " + CompletionChunk(id="cmpl-1", model=request.model, choices=[
                        CompletionChunkChoice(index=0, text=piece, finish_reason="length")]).model_dump_json() + "\n\n"
                yield "data: [DONE]\n\n"
            return gen()
        return Completion(id="cmpl-1", model=request.model, object="text_completion", created=1, choices=[CompletionChoice(index=0, text=text, finish_reason="length")],
                          usage=UsageInfo(prompt_tokens=4, completion_tokens=7, total_tokens=11))


@pytest.fixture(scope="module")
def client():
    server = ModelServer()
    app = server.create_application([EchoModel("echo"), DummyOpenAI("gpt")])
    with TestClient(app) as c:
        yield c


def test_health_and_metadata(client):
    assert client.get("/").json() == {"status": "alive"}
    assert client.get("/v2/health/ready").json() == {"ready": True}
    assert client.get("/v1/models/echo").json() == {"name": "echo", "ready": True}
    assert client.get("/v1/models/nope").status_code == 404
    assert "request_predict_seconds" in client.get("/metrics").text


def test_v1_predict(client):
    r = client.post("/v1/models/echo:predict", json={"instances": [[1, 2, 3], [4, 5, 6]]})
    assert r.status_code == 200 and r.json() == {"predictions": [[1, 2, 3], [4, 5, 6]]}
    assert client.post("/v1/models/echo:predict", json={"instances": 3}).status_code == 400
    assert client.post("/v1/models/echo:predict", content=b"{not json").status_code == 400
    assert client.post("/v1/models/missing:predict", json={"instances": [1]}).status_code == 404


def test_v2_infer_json_and_binary(client):
    body = {"id": "42", "inputs": [{"name": "input_ids", "shape": [2, 3], "datatype": "INT64", "data": [1, 2, 3, 4, 5, 6]}]}
    r = client.post("/v2/models/echo/infer", json=body)
    assert r.status_code == 200
    out = r.json()
    assert out["id"] == "42" and out["outputs"][0]["data"] == [1, 2, 3, 4, 5, 6] and out["outputs"][0]["shape"] == [2, 3]
    # binary extension: JSON header + raw little-endian tensor, Inference-Header-Content-Length (infer_type.py:593-666)
    arr = np.arange(12, dtype=np.int64).reshape(3, 4)
    hdr = json.dumps({"inputs": [{"name": "input_ids", "shape": [3, 4], "datatype": "INT64",
                                  "parameters": {"binary_data_size": arr.nbytes}}],
                      "parameters": {"binary_data_output": True}}).encode()
    r = client.post("/v2/models/echo/infer", content=hdr + arr.tobytes(),
                    headers={"Inference-Header-Content-Length": str(len(hdr)), "Content-Type": "application/octet-stream"})
    assert r.status_code == 200
    n = int(r.headers["inference-header-content-length"])
    meta = json.loads(r.content[:n])
    back = np.frombuffer(r.content[n:], dtype=np.int64).reshape(meta["outputs"][0]["shape"])
    assert np.array_equal(back, arr) and meta["outputs"][0]["parameters"]["binary_data_size"] == arr.nbytes


def test_infer_type_roundtrip_fp16_and_bytes():
    x = InferInput("x", [2, 2], "FP16")
    x.set_data_from_numpy(np.array([[1, 2], [3, 4]], dtype=np.float16))
    t = InferInput("t", [2], "BYTES")
    t.set_data_from_numpy(np.array([b"hello", "wörld".encode()], dtype=np.object_))
    raw, n = InferRequest("m", [x, t], request_id="r1").to_rest()
    back = InferRequest.from_bytes(raw, n, "m")
    assert back.id == "r1"
    assert np.array_equal(back.inputs[0].as_numpy(), x.as_numpy())
    assert back.inputs[1].as_string() == ["hello", "wörld"]
    with pytest.raises(InvalidInput):   # FP16 via JSON is rejected (infer_type.py:625-629)
        InferRequest.from_dict({"inputs": [{"name": "x", "shape": [1], "datatype": "FP16", "data": [1.0]}]}, "m")
    with pytest.raises(InvalidInput):
        InferRequest.from_dict({"inputs": [{"name": "x", "shape": [1], "datatype": "FP32"}]}, "m")


def test_zero_copy_binary_input():
    arr = np.arange(8, dtype=np.int64)
    hdr = json.dumps({"inputs": [{"name": "a", "shape": [8], "datatype": "INT64", "parameters": {"binary_data_size": 64}}]}).encode()
    buf = hdr + arr.tobytes()
    req = InferRequest.from_bytes(buf, len(hdr), "m")
    view = req.inputs[0].as_numpy()
    assert np.array_equal(view, arr) and not view.flags.owndata    # a view on the request buffer, not a list round trip


def test_openai_completion_and_errors(client):
    r = client.post("/openai/v1/completions", json={"model": "gpt", "prompt": "hi"})
    assert r.status_code == 200
    j = r.json()
    assert j["object"] == "text_completion" and j["choices"][0] == {"index": 0, "text": "echo:hi", "finish_reason": "length"}
    assert j["usage"] == {"prompt_tokens": 4, "total_tokens": 11, "completion_tokens": 7}
    r = client.post("/openai/v1/completions", json={"model": "gpt", "prompt": "boom"})
    assert r.status_code == 500
    assert r.json()["error"] == {"code": "500", "message": "'frequency_penalty' is not supported", "param": "", "type": "OpenAIError"}
    assert client.post("/openai/v1/completions", json={"model": "nope", "prompt": "x"}).status_code == 404
    assert client.post("/openai/v1/completions", json={"model": "echo", "prompt": "x"}).status_code == 400  # not an OpenAI model
    assert [m["id"] for m in client.get("/openai/v1/models").json()["data"]] == ["gpt"]


def test_openai_streaming_sse_and_chat_mapping(client):
    with client.stream("POST", "/openai/v1/completions", json={"model": "gpt", "prompt": "hi", "stream": True}) as r:
        assert r.headers["content-type"].startswith("text/event-stream")
        lines = [l for l in r.iter_lines() if l]
    assert lines[-1] == "data: [DONE]"
    chunks = [json.loads(l[len("data: "):]) for l in lines[:-1]]
    assert "".join(c["choices"][0]["text"] for c in chunks) == "echo:hi"
    assert len({c["id"] for c in chunks}) == 1 and all(c["choices"][0]["finish_reason"] == "length" for c in chunks)
    chat = {"model": "gpt", "messages": [{"role": "user", "content": "yo"}]}
    j = client.post("/openai/v1/chat/completions", json=chat).json()
    assert j["object"] == "chat.completion" and j["choices"][0]["message"] == {"role": "assistant", "content": "echo:<|user|>yo\n"}   # unset fields are excluded on the wire
    with client.stream("POST", "/openai/v1/chat/completions", json={**chat, "stream": True}) as r:
        lines = [l for l in r.iter_lines() if l]
    assert lines[-1] == "data: [DONE]"
    deltas = [json.loads(l[6:])["choices"][0]["delta"]["content"] for l in lines[:-1]]
    assert "".join(deltas) == "echo:<|user|>yo\n"
    assert client.post("/openai/v1/chat/completions", json={**chat, "n": 2}).status_code == 400


def test_completion_request_defaults_match_vllm_type():
    """max_tokens=16, stop=[], temperature=None, top_p=None, n=1, echo=False, stream=False."""
    r = CompletionRequest(prompt="x")
    assert (r.max_tokens, r.stop, r.temperature, r.top_p, r.n, r.echo, r.stream) == (16, [], None, None, 1, False, False)
