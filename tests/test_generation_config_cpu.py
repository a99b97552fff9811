"""Request -> generation parameters, host logic only (generative_model.py:388-402 + transformers' default merging)."""
from types import SimpleNamespace

from kserve_b200.generative_model import B200GenerativeModel


def _model(defaults=None):
    m = B200GenerativeModel.__new__(B200GenerativeModel)
    m.generation_defaults = dict(defaults or {})
    m._seed_counter = 0
    return m


def _req(**kw):
    base = dict(presence_penalty=None, temperature=None, top_p=None, seed=None)
    base.update(kw)
    return SimpleNamespace(**base)


def test_greedy_unless_the_checkpoint_enables_sampling():
    m = _model()
    assert m.build_generation_config(_req()) == {}
    assert m.build_generation_config(_req(temperature=0.7, top_p=0.9)) == {}          # do_sample is never set by the reference (q9)
    assert m.build_generation_config(_req(presence_penalty=1.3)) == {"repetition_penalty": 1.3}   # q8
    assert m.build_generation_config(_req(presence_penalty=0)) == {}
    assert m.build_generation_config(_req(presence_penalty=-0.5)) == {}


def test_checkpoint_defaults_fill_the_none_fields():
    m = _model({"do_sample": True, "temperature": 0.6, "top_p": 0.9})
    g = m.build_generation_config(_req(seed=11))
    assert g == {"do_sample": True, "temperature": 0.6, "top_p": 0.9, "top_k": 50, "seed": 11}
    g = m.build_generation_config(_req(temperature=1.2, top_p=0.5, seed=3))
    assert (g["temperature"], g["top_p"], g["top_k"]) == (1.2, 0.5, 50)
    assert m.build_generation_config(_req(temperature=0)) == {}                        # temperature 0 -> greedy
    a = m.build_generation_config(_req())["seed"]
    b = m.build_generation_config(_req())["seed"]
    assert a != b                                                                      # unseeded requests do not repeat
    m2 = _model({"do_sample": True, "top_k": 0, "repetition_penalty": 1.1})
    g = m2.build_generation_config(_req(seed=1))
    assert g["top_k"] == 1024 and g["repetition_penalty"] == 1.1 and g["temperature"] == 1.0 and g["top_p"] == 1.0


def test_chat_template_is_required_like_the_reference():
    """generative_model.py:495-500: no template on the request and none in the tokenizer -> OpenAIError with this text"""
    import pytest
    from kserve_b200.kserve_api.protocol.rest.openai.errors import OpenAIError
    from kserve_b200.kserve_api.protocol.rest.openai.types import ChatCompletionRequest
    m = _model()
    m._tokenizer = SimpleNamespace(chat_template=None)
    req = ChatCompletionRequest.model_validate({"model": "m", "messages": [{"role": "user", "content": "hi"}]})
    with pytest.raises(OpenAIError) as e:
        m.apply_chat_template(req)
    assert "default chat template is no longer allowed" in str(e.value)
    calls = {}
    m._tokenizer = SimpleNamespace(chat_template="{{x}}", apply_chat_template=lambda **kw: calls.update(kw) or "PROMPT")
    assert m.apply_chat_template(req).prompt == "PROMPT"
    assert calls["tokenize"] is False and calls["conversation"] == [{"role": "user", "content": "hi"}]
