"""Checkpoint-config handling of load(): rope tables bit-equal to transformers' LlamaRotaryEmbedding for every supported
rope type, refusal of the unsupported ones, the context-length derivation of the reference
(python/huggingfaceserver/huggingfaceserver/utils.py:28-159) and the device-memory plan."""
import pytest
import torch

from kserve_b200.utils import (get_and_verify_max_len, hf_rope_tables, plan_memory, rope_inv_freq, rope_parameters)

LLAMA31_ROPE = {"factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0, "original_max_position_embeddings": 8192,
                "rope_type": "llama3"}


def _hf_tables(cfg_kwargs, max_pos):
    """cos/sin from transformers' own rotary module for a bf16 activation tensor"""
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding
    cfg = LlamaConfig(hidden_size=512, num_attention_heads=4, num_key_value_heads=2, head_dim=128, intermediate_size=256,
                      num_hidden_layers=1, vocab_size=64, **cfg_kwargs)
    rot = LlamaRotaryEmbedding(cfg)
    x = torch.zeros(1, max_pos, 8, dtype=torch.bfloat16)
    cos, sin = rot(x, torch.arange(max_pos)[None])
    return cos[0, :, :64], sin[0, :, :64]      # the two halves of the 128 columns are copies of each other


@pytest.mark.parametrize("name,cfg", [
    ("default", dict(rope_theta=500000.0, max_position_embeddings=8192)),
    ("linear", dict(rope_theta=10000.0, max_position_embeddings=4096, rope_scaling={"rope_type": "linear", "factor": 4.0})),
    ("llama3", dict(rope_theta=500000.0, max_position_embeddings=131072, rope_scaling=dict(LLAMA31_ROPE))),
    ("llama3_old_type_key", dict(rope_theta=500000.0, max_position_embeddings=131072,
                                 rope_scaling={**{k: v for k, v in LLAMA31_ROPE.items() if k != "rope_type"}, "type": "llama3"})),
])
def test_rope_tables_bit_equal_to_transformers(name, cfg):
    n = 9000 if name.startswith("llama3") else 3000
    cos, sin = hf_rope_tables(dict(cfg), 128, n)
    hf_cfg = dict(cfg)
    rs = hf_cfg.pop("rope_scaling", None)
    if rs is not None:
        rs = dict(rs)
        if "type" in rs and "rope_type" not in rs:
            rs["rope_type"] = rs.pop("type")
        hf_cfg["rope_parameters"] = {**rs, "rope_theta": hf_cfg.pop("rope_theta")}
    hc, hs = _hf_tables(hf_cfg, n)
    assert torch.equal(cos, hc) and torch.equal(sin, hs)
    if name != "default":      # and the scaling actually changes something
        c0, _ = hf_rope_tables({"rope_theta": cfg["rope_theta"]}, 128, n)
        assert not torch.equal(c0, cos)


def test_unsupported_rope_types_are_refused():
    for kind in ("yarn", "dynamic", "longrope"):
        with pytest.raises(ValueError, match="not supported"):
            rope_inv_freq({"rope_theta": 1e4, "rope_scaling": {"rope_type": kind, "factor": 2.0}}, 128)
    assert rope_parameters({"rope_theta": 5.0})["rope_type"] == "default"


def test_get_and_verify_max_len_matches_reference_rules(monkeypatch):
    # smallest of the known keys; nothing found -> 2048 (the bloom case of tests/test_model.py:356)
    assert get_and_verify_max_len({"max_position_embeddings": 8192, "n_positions": 1024}, None) == 1024
    assert get_and_verify_max_len({}, None) == 2048
    assert get_and_verify_max_len({}, 777) == 777
    # rope scaling: llama3 keeps max_position_embeddings, linear multiplies by the factor, yarn starts from the original length
    assert get_and_verify_max_len({"max_position_embeddings": 131072, "rope_scaling": dict(LLAMA31_ROPE)}, None) == 131072
    assert get_and_verify_max_len({"max_position_embeddings": 4096, "rope_scaling": {"rope_type": "linear", "factor": 4.0}}, None) == 16384
    assert get_and_verify_max_len({"max_position_embeddings": 32768, "rope_scaling": {
        "rope_type": "yarn", "factor": 4.0, "original_max_position_embeddings": 8192}}, None) == 32768
    # a user value below the derived one wins; above it is an error unless the env override is set
    assert get_and_verify_max_len({"max_position_embeddings": 8192}, 512) == 512
    with pytest.raises(ValueError, match=r"User-specified max_model_len \(9000\) is greater than the derived max_model_len "
                                         r"\(max_position_embeddings=8192"):
        get_and_verify_max_len({"max_position_embeddings": 8192}, 9000)
    monkeypatch.setenv("ALLOW_LONG_MAX_MODEL_LEN", "1")
    assert get_and_verify_max_len({"max_position_embeddings": 8192}, 9000) == 9000
    monkeypatch.delenv("ALLOW_LONG_MAX_MODEL_LEN")
    # model_max_length may exceed the derived length
    assert get_and_verify_max_len({"max_position_embeddings": 2048, "model_max_length": 4096}, 4096) == 4096
    assert get_and_verify_max_len({"max_position_embeddings": 8192, "sliding_window": 4096}, None, True, 4096) == 4096


def test_plan_memory_caps_prefill_buffers_and_kv_pool():
    llama31 = dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                   num_key_value_heads=8, head_dim=128)
    free = 178 * 2**30
    cap_t, pages = plan_memory(llama31, 128256, 32, 131072, 1, free)
    assert cap_t == 131072                                   # one full-length prompt fits; not 32 x 131072 rows
    page_bytes = 32 * 2 * 8 * 64 * 128 * 2
    assert 2048 <= pages < 32 * 2048 and pages * page_bytes < free - 16 * 2**30
    # the bench configuration gets everything it asks for
    cap_t, pages = plan_memory(llama31, 128256, 32, 1152, 1, free)
    assert (cap_t, pages) == (32768, 32 * 18)
    with pytest.raises(ValueError, match="not enough device memory"):
        plan_memory(llama31, 128256, 32, 131072, 1, 20 * 2**30)
    # tensor parallel shards weights and KV heads
    assert plan_memory(llama31, 128256, 32, 131072, 8, free)[1] > 4 * pages
