"""DataPlane behaviour, restating the non-CloudEvent scenarios of the reference's
python/kserve/test/test_dataplane.py:83-165 and :425-455: registry lookups and their error text, liveness / readiness
semantics, metadata shapes, decode -> infer / explain -> encode, and the type gate that keeps OpenAI-only models off
the V1/V2 inference routes."""
import asyncio

import pytest

from kserve_b200.kserve_api import Model
from kserve_b200.kserve_api.errors import ModelNotFound
from kserve_b200.kserve_api.model_repository import ModelRepository
from kserve_b200.kserve_api.protocol.dataplane import DataPlane
from kserve_b200.kserve_api.protocol.rest.openai.openai_model import OpenAIGenerativeModel


class Dummy(Model):
    def __init__(self, name):
        super().__init__(name)
        self.ready = False

    def load(self):
        self.ready = True

    async def predict(self, request, headers=None):
        return {"predictions": request["inputs"] if "inputs" in request else request["instances"]}

    async def explain(self, request, headers=None):
        return {"predictions": request["inputs"] if "inputs" in request else request["instances"]}


def _dp_with_model(name="TestModel"):
    dp = DataPlane(model_registry=ModelRepository())
    m = Dummy(name)
    m.load()
    dp._model_registry.update(m)
    return dp


def test_registry_lookup_and_error_text():
    dp = DataPlane(model_registry=ModelRepository())
    with pytest.raises(ModelNotFound) as e:
        dp.get_model_from_registry("FakeModel")
    assert e.value.reason == "Model with name FakeModel does not exist."
    dp = _dp_with_model("Model")
    assert dp.get_model_from_registry("Model").name == "Model"


def test_liveness_and_readiness_semantics():
    assert asyncio.run(DataPlane.live()) == {"status": "alive"}
    dp = _dp_with_model()
    dp._model_registry.update(Dummy("NotReadyModel"))          # registered but never loaded
    assert asyncio.run(dp.ready()) is True                     # server readiness ignores the models' readiness
    assert asyncio.run(dp.model_ready("TestModel")) is True
    assert asyncio.run(dp.model_ready("NotReadyModel")) is False


def test_metadata_shapes():
    dp = _dp_with_model()
    md = dp.metadata()
    assert md["name"] == "kserve" and md["extensions"] == ["model_repository_extension"] and isinstance(md["version"], str)
    assert asyncio.run(dp.model_metadata("TestModel")) == {"name": "TestModel", "platform": "", "inputs": [], "outputs": []}


def test_decode_infer_explain_encode():
    dp = _dp_with_model()
    for call in (dp.infer, dp.explain):
        req, attrs = dp.decode(b'{"instances":[[1,2]]}', None)
        resp, headers = asyncio.run(call("TestModel", req))
        assert dp.encode("TestModel", resp, headers, attrs) == ({"predictions": [[1, 2]]}, {})


def test_openai_only_models_are_not_inference_models():
    class OnlyOpenAI(OpenAIGenerativeModel):
        async def create_completion(self, params):
            pass

        async def create_chat_completion(self, params):
            pass
    repo = ModelRepository()
    repo.update(OnlyOpenAI("TestModel"))
    with pytest.raises(ValueError) as e:
        asyncio.run(DataPlane(model_registry=repo).infer(model_name="TestModel", request={}))
    assert e.value.args[0] == "Model of type OnlyOpenAI does not support inference"


def test_model_repository_scenarios():
    """python/kserve/test/test_model_repository.py:55-135: aliases, OpenAI-only models, readiness of unknown / unloaded
    models (an OpenAI model without a load step counts as ready)."""
    repo = ModelRepository()
    m = Model(name="kserve-model")
    repo.update(m)
    assert repo.get_model("kserve-model") is m and m.name == "kserve-model"
    repo.update(m, name="additional-model-name")
    assert repo.get_model("additional-model-name") is m
    assert asyncio.run(repo.is_model_ready("none-model")) is False
    assert asyncio.run(repo.is_model_ready("kserve-model")) is False
    m.load()
    assert asyncio.run(repo.is_model_ready("kserve-model")) is True

    class OnlyOpenAI(OpenAIGenerativeModel):
        async def create_completion(self, params):
            pass

        async def create_chat_completion(self, params):
            pass
    o = OnlyOpenAI(name="openai-model")
    repo.update(o)
    assert isinstance(repo.get_model("openai-model"), OpenAIGenerativeModel)
    assert asyncio.run(repo.is_model_ready("openai-model")) is True


def test_predict_input_and_response_helpers():
    """kserve/utils/utils.py:149-254: the payload forms custom models consume / return"""
    import numpy as np
    import pandas as pd
    from kserve_b200.kserve_api.protocol.infer_type import (InferInput, InferRequest, get_predict_input, get_predict_response)
    assert get_predict_input({"instances": [[1, 2], [3, 4]]}).tolist() == [[1, 2], [3, 4]]
    assert get_predict_input({"inputs": ["a", "b"]}) == ["a", "b"]
    assert get_predict_input({"instances": []}).shape == (0,)
    df = get_predict_input({"instances": [{"a": [1], "b": [2]}, {"a": [3], "b": [4]}]})
    assert isinstance(df, pd.DataFrame) and df["a"].tolist() == [1, 3] and df["b"].tolist() == [2, 4]
    req = InferRequest(model_name="m", request_id="1", infer_inputs=[InferInput("x", [2], "INT32", data=[5, 6]),
                                                                     InferInput("s", [2], "BYTES", data=["u", "v"])])
    assert get_predict_input(req).tolist() == [5, 6]
    req.parameters = {"content_type": "pd"}
    df = get_predict_input(req)
    assert list(df.columns) == ["x", "s"] and df["s"].tolist() == ["u", "v"]
    assert get_predict_input(InferRequest(model_name="m", infer_inputs=[InferInput("s", [2], "BYTES", data=["u", "v"])])) == ["u", "v"]
    # responses
    assert get_predict_response({"instances": []}, np.array([[1, 2]]), "m") == {"predictions": [[1, 2]]}
    assert get_predict_response({"instances": []}, pd.DataFrame({"a": [1, 2]}), "m") == {"predictions": [{"a": 1}, {"a": 2}]}
    r = get_predict_response(req, pd.DataFrame({"p": np.array([1.5, 2.5], dtype=np.float32), "q": np.array([1, 2], dtype=np.int64)}), "m")
    assert [(o.name, o.datatype, o.shape) for o in r.outputs] == [("p", "FP32", [2]), ("q", "INT64", [2])] and r.id == "1"
    r = get_predict_response(req, ["cat", "dog"], "m")
    assert (r.outputs[0].name, r.outputs[0].datatype, r.outputs[0].shape) == ("output-0", "BYTES", [2])
    assert r.to_rest()[0]["outputs"][0]["data"] == ["cat", "dog"]
    r = get_predict_response(req, [[1, 2], [3, 4]], "m")
    assert r.outputs[0].shape == [2, 2] and r.outputs[0].datatype == "INT64"
