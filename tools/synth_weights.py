"""Deterministic synthetic Llama checkpoints (test / bench infrastructure; re-exported as oracle.weights).

There is no network and no checkpoint on disk, so both the CPU oracle and the
CUDA engine are fed the same seeded tensors.  Each tensor gets its own CPU
``torch.Generator`` (seed derived from the tensor's index in a canonical
order) so that tensors can be produced one at a time without holding a whole
fp32 model in memory, and so that the bytes are identical on every box.

Names follow the HF Llama ``state_dict`` (what the reference's
``AutoModelForCausalLM.from_pretrained`` would load,
python/huggingfaceserver/huggingfaceserver/generative_model.py:249-254).
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, Tuple

import torch

# Named model shapes used by tests / bench.  ``llama3_8b`` is BASELINE.json's
# flagship; the tiny ones exercise GQA group sizes 2 and 4 and a vocabulary
# that is not a multiple of any tile size.
CONFIGS: Dict[str, dict] = {
    "tiny_g2": dict(vocab_size=1031, hidden_size=512, intermediate_size=1024,
                    num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                    head_dim=128, max_position_embeddings=2048, rms_norm_eps=1e-5,
                    rope_theta=500000.0),
    "tiny_g4": dict(vocab_size=2048, hidden_size=1024, intermediate_size=2816,
                    num_hidden_layers=3, num_attention_heads=8, num_key_value_heads=2,
                    head_dim=128, max_position_embeddings=2048, rms_norm_eps=1e-5,
                    rope_theta=500000.0),
    "llama3_8b": dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336,
                      num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
                      head_dim=128, max_position_embeddings=8192, rms_norm_eps=1e-5,
                      rope_theta=500000.0),
    # same layer shapes as llama3_8b, 2 layers: the largest thing the CPU oracle
    # can run in seconds while still hitting every full-size GEMM / vocab shape.
    "llama3_8b_2l": dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336,
                         num_hidden_layers=2, num_attention_heads=32, num_key_value_heads=8,
                         head_dim=128, max_position_embeddings=8192, rms_norm_eps=1e-5,
                         rope_theta=500000.0),
}


MOE_CONFIGS: Dict[str, dict] = {
    # Mixtral-style sparse MoE (BASELINE.json configs[3]) at test size: 4 experts, top-2
    "tiny_moe": dict(vocab_size=1031, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                     num_attention_heads=4, num_key_value_heads=2, head_dim=128, max_position_embeddings=2048,
                     rms_norm_eps=1e-5, rope_theta=1000000.0, num_local_experts=4, num_experts_per_tok=2),
    "tiny_moe8": dict(vocab_size=2048, hidden_size=1024, intermediate_size=1536, num_hidden_layers=2,
                      num_attention_heads=8, num_key_value_heads=2, head_dim=128, max_position_embeddings=2048,
                      rms_norm_eps=1e-5, rope_theta=1000000.0, num_local_experts=8, num_experts_per_tok=2),
}
CONFIGS.update(MOE_CONFIGS)

# 8 KV heads: the smallest shapes every tensor-parallel degree of BASELINE configs[2] (TP = 2 / 4 / 8) can shard
# (TP=8 -> one KV head and two query heads per GPU); the vocabulary is odd on purpose (uneven vocab-parallel shards).
CONFIGS["tiny_kv8"] = dict(vocab_size=4099, hidden_size=1024, intermediate_size=2048, num_hidden_layers=2,
                           num_attention_heads=16, num_key_value_heads=8, head_dim=128, max_position_embeddings=2048,
                           rms_norm_eps=1e-5, rope_theta=500000.0)
# Llama-3.1-style rope scaling (rope_type "llama3"): the tables differ from the default ones at every position
CONFIGS["tiny_g2_rope3"] = dict(CONFIGS["tiny_g2"], max_position_embeddings=4096,
                                rope_scaling=dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                                  original_max_position_embeddings=256))

# "Peaked" variants: lm_head row v = embed row perm[v].  The final hidden state still carries the current token's
# embedding (the residual stream starts from it), so one logit towers over the rest the way a trained model's does and
# the greedy choice is decisive at (nearly) every step: greedy ids can then be compared EXACTLY, not just where a
# random-weight near-tie happens to be wide enough.  All other tensors are those of the base config, so the layers are
# exercised identically and the full logits are still compared element-wise (tests/helpers.py: logits_tol_elementwise).
for _base in ("tiny_g2", "tiny_g4", "tiny_kv8", "llama3_8b_2l"):
    CONFIGS[_base + "_peaked"] = dict(CONFIGS[_base], _peaked=True)


def hf_config_dict(cfg: dict) -> dict:
    """the keys a transformers config understands (drops this module's private `_...` markers)"""
    return {k: v for k, v in cfg.items() if not k.startswith("_")}


def moe_tensor_specs(cfg: dict) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """HF MixtralForCausalLM state_dict (transformers 5.x fused-expert layout)."""
    H, I, V, E = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"], cfg["num_local_experts"]
    nh, nkv, d = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"]
    yield "model.embed_tokens.weight", (V, H), "embed"
    for i in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{i}."
        yield p + "input_layernorm.weight", (H,), "norm"
        yield p + "self_attn.q_proj.weight", (nh * d, H), "linear"
        yield p + "self_attn.k_proj.weight", (nkv * d, H), "linear"
        yield p + "self_attn.v_proj.weight", (nkv * d, H), "linear"
        yield p + "self_attn.o_proj.weight", (H, nh * d), "linear"
        yield p + "post_attention_layernorm.weight", (H,), "norm"
        yield p + "mlp.gate.weight", (E, H), "router"
        yield p + "mlp.experts.gate_up_proj", (E, 2 * I, H), "linear"
        yield p + "mlp.experts.down_proj", (E, H, I), "linear"
    yield "model.norm.weight", (H,), "norm"
    yield "lm_head.weight", (V, H), "linear"


def tensor_specs(cfg: dict) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (name, shape, kind) in canonical order. kind in {embed, linear, norm}."""
    H, I, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"]
    nh, nkv, d = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"]
    if "num_local_experts" in cfg:
        yield from moe_tensor_specs(cfg)
        return
    yield "model.embed_tokens.weight", (V, H), "embed"
    for i in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{i}."
        yield p + "input_layernorm.weight", (H,), "norm"
        yield p + "self_attn.q_proj.weight", (nh * d, H), "linear"
        yield p + "self_attn.k_proj.weight", (nkv * d, H), "linear"
        yield p + "self_attn.v_proj.weight", (nkv * d, H), "linear"
        yield p + "self_attn.o_proj.weight", (H, nh * d), "linear"
        yield p + "post_attention_layernorm.weight", (H,), "norm"
        yield p + "mlp.gate_proj.weight", (I, H), "linear"
        yield p + "mlp.up_proj.weight", (I, H), "linear"
        yield p + "mlp.down_proj.weight", (H, I), "linear"
    yield "model.norm.weight", (H,), "norm"
    yield "lm_head.weight", (V, H), "linear"


def synth_tensor(index: int, shape, kind: str, seed: int, dtype=torch.bfloat16) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed * 1000003 + index * 7919 + 17)
    if kind == "norm":
        t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
    elif kind == "embed":
        t = torch.randn(shape, generator=g, dtype=torch.float32)
    elif kind == "router":  # moderately peaky routing: the 2nd expert carries real weight, 2nd/3rd margins mostly above bf16 noise
        t = torch.randn(shape, generator=g, dtype=torch.float32) * (2.5 / math.sqrt(shape[-1]))
    else:  # linear: variance preserving, so attention / logits are not near-uniform
        t = torch.randn(shape, generator=g, dtype=torch.float32) * (1.0 / math.sqrt(shape[-1]))
    return t.to(dtype)


def peaked_perm(vocab: int, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed * 7727 + 4241)
    return torch.randperm(vocab, generator=g)


def iter_state_dict(cfg: dict, seed: int = 0, dtype=torch.bfloat16):
    """Yield (name, tensor) one tensor at a time (bounded memory)."""
    for idx, (name, shape, kind) in enumerate(tensor_specs(cfg)):
        if name == "lm_head.weight" and cfg.get("_peaked"):
            embed = synth_tensor(0, shape, "embed", seed, dtype)      # index 0 == model.embed_tokens.weight
            yield name, embed[peaked_perm(shape[0], seed)].contiguous()
            continue
        yield name, synth_tensor(idx, shape, kind, seed, dtype)


def synth_state_dict(cfg: dict, seed: int = 0, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    return dict(iter_state_dict(cfg, seed, dtype))


def checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Cheap fingerprint recorded in golden files to prove both sides loaded the same bytes."""
    s = 0.0
    third = "model.layers.0.mlp.down_proj.weight"
    if third not in sd:   # MoE checkpoints have fused expert tensors instead
        third = "model.layers.0.self_attn.o_proj.weight"
    for name in ("model.embed_tokens.weight", "lm_head.weight", third):
        t = sd[name]
        s += float(t[: 64].float().sum()) + float(t[-1].float().abs().sum())
    return s
