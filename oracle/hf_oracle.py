"""CPU oracle: the reference's HF-transformers generate path, restated.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every block cites the
reference lines it follows; paths are relative to
/root/reference/python/huggingfaceserver/huggingfaceserver/ unless noted.

The arithmetic itself is *not* restated: exactly like the reference
(generative_model.py:328 ``outputs = self._model.generate(**kwargs)``) it is
delegated to ``transformers`` on the CPU backend.  What is restated is the
wrapper that decides what goes into and comes out of that call.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Union

import torch
from transformers import (
    GenerationConfig,
    LlamaConfig,
    LlamaForCausalLM,
    StoppingCriteria,
    StoppingCriteriaList,
)


class OracleError(Exception):
    """Stands in for kserve's OpenAIError (protocol/rest/openai/errors.py:22-36)."""


class StopSequenceStoppingCriteria(StoppingCriteria):
    """stop_sequence_stopping_criteria.py:21-48 — batch-wide: any row matching stops all rows."""

    def __init__(self, input_length: int, stop_sequences: List[torch.Tensor]):
        self.input_length = input_length
        self.stop_sequences = stop_sequences
        self.triggered = False

    def __call__(self, input_ids, scores, **kwargs) -> bool:
        for seq in self.stop_sequences:
            if seq.shape[-1] > input_ids.shape[-1] - self.input_length:  # :41-42
                continue
            if torch.any(torch.all(input_ids[:, -len(seq):] == seq, dim=1)):  # :45
                self.triggered = True
                return True
        return False


def build_llama(cfg: dict, state_dict: Dict[str, torch.Tensor], dtype=torch.bfloat16,
                eos_token_id=None) -> LlamaForCausalLM:
    """from_config + load seeded weights (the reference uses from_pretrained, :249-254)."""
    cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}      # tools/synth_weights.py private markers
    if cfg.get("rope_scaling"):        # transformers 5.x spelling of the 4.x `rope_scaling` + `rope_theta` pair
        cfg["rope_parameters"] = {**cfg.pop("rope_scaling"), "rope_theta": cfg.pop("rope_theta")}
    hf_cfg = LlamaConfig(**cfg, tie_word_embeddings=False, eos_token_id=eos_token_id,
                         bos_token_id=None, pad_token_id=None, attention_bias=False, mlp_bias=False)
    with torch.device("meta"):
        model = LlamaForCausalLM(hf_cfg)
    model = model.to_empty(device="cpu").to(dtype)
    missing, unexpected = model.load_state_dict(state_dict, strict=False, assign=True)
    assert not unexpected, unexpected
    # rotary inv_freq is a non-persistent buffer: to_empty() left it uninitialised, and it must stay
    # fp32 as it does under the reference's from_pretrained(torch_dtype=...) (modeling_llama.py:84-90)
    n_fixed = 0
    for mod in model.modules():
        if hasattr(mod, "inv_freq"):
            from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
            init = (type(mod).compute_default_rope_parameters if mod.rope_type == "default"
                    else ROPE_INIT_FUNCTIONS[mod.rope_type])      # what LlamaRotaryEmbedding.__init__ selects
            inv_freq, scaling = init(mod.config, "cpu")
            mod.register_buffer("inv_freq", inv_freq.float(), persistent=False)
            mod.register_buffer("original_inv_freq", inv_freq.float().clone(), persistent=False)
            mod.attention_scaling = scaling
            n_fixed += 1
    assert n_fixed >= 1, "no rotary module found"
    bad = [m for m in missing if "inv_freq" not in m]
    assert not bad, bad
    model.eval()  # :255
    return model


def build_mixtral(cfg: dict, state_dict: Dict[str, torch.Tensor], dtype=torch.bfloat16):
    """MixtralForCausalLM from config + seeded weights (BASELINE configs[3] architecture)."""
    from transformers import MixtralConfig, MixtralForCausalLM
    hf_cfg = MixtralConfig(**cfg, tie_word_embeddings=False, eos_token_id=None, bos_token_id=None, pad_token_id=None,
                           sliding_window=None, router_jitter_noise=0.0)
    with torch.device("meta"):
        model = MixtralForCausalLM(hf_cfg)
    model = model.to_empty(device="cpu").to(dtype)
    missing, unexpected = model.load_state_dict(state_dict, strict=False, assign=True)
    assert not unexpected, unexpected
    n_fixed = 0
    for mod in model.modules():
        if hasattr(mod, "inv_freq"):
            inv_freq, scaling = type(mod).compute_default_rope_parameters(mod.config, "cpu")
            mod.register_buffer("inv_freq", inv_freq.float(), persistent=False)
            mod.register_buffer("original_inv_freq", inv_freq.float().clone(), persistent=False)
            mod.attention_scaling = scaling
            n_fixed += 1
    assert n_fixed >= 1
    assert not [m for m in missing if "inv_freq" not in m], missing
    model.eval()
    return model


def build_model(cfg: dict, state_dict, dtype=torch.bfloat16):
    return build_mixtral(cfg, state_dict, dtype) if "num_local_experts" in cfg else build_llama(cfg, state_dict, dtype)


@dataclass
class OracleResult:
    output_ids: torch.Tensor            # [B, S+T] what generate() returned (prompt echoed)
    texts: Optional[List[str]]          # batch_decode(outputs[:, start:]) or None (no tokenizer)
    finish_reason: str
    prompt_tokens: int
    completion_tokens: int
    step_logits: Optional[List[torch.Tensor]] = None   # per step fp32 [B, V] (processed == raw for greedy)
    seconds: float = 0.0
    prefill_seconds: float = 0.0
    extra: Dict[str, Any] = field(default_factory=dict)


class OracleGenerativeModel:
    """Restated HuggingfaceGenerativeModel (generative_model.py:143-646), decoder-only path."""

    def __init__(self, model, tokenizer=None, pad_token_id: Optional[int] = None,
                 max_length: Optional[int] = None):
        self._model = model
        self._tokenizer = tokenizer
        if tokenizer is not None:
            # :225-231 decoder-only => left padding; :256-265 fallback [PAD]
            tokenizer.padding_side = "left"
            if not tokenizer.pad_token:
                tokenizer.add_special_tokens({"pad_token": "[PAD]"})
                model.resize_token_embeddings(len(tokenizer))
            pad_token_id = tokenizer.pad_token_id
        self.pad_token_id = pad_token_id
        self.max_length = max_length or model.config.max_position_embeddings  # utils.py:28-159 (simplified)

    # generative_model.py:388-402 — nothing else is set; do_sample stays unset => greedy
    def build_generation_config(self, max_tokens, top_p=None, temperature=None,
                                presence_penalty=None) -> GenerationConfig:
        kwargs = {"max_new_tokens": max_tokens, "top_p": top_p, "temperature": temperature,
                  "pad_token_id": self.pad_token_id}
        if presence_penalty and presence_penalty > 0:
            kwargs["repetition_penalty"] = presence_penalty
        return GenerationConfig(**kwargs)

    @torch.no_grad()  # :341
    def create_completion(self, prompt: Union[str, List[str], List[int], List[List[int]]],
                          max_tokens: Optional[int] = 16, stop: Union[None, str, List[str], List[List[int]]] = (),
                          echo: bool = False, temperature=None, top_p=None,
                          presence_penalty=None, frequency_penalty=None, n=None,
                          want_logits: bool = False) -> OracleResult:
        if prompt is None:
            raise OracleError("prompt is required")  # :542-543
        # :546-551 prompt normalisation
        prompts = prompt if isinstance(prompt, list) and not isinstance(prompt[0], int) else [prompt]
        if isinstance(prompts[0][0], int):
            inputs = {"input_ids": torch.tensor(prompts, dtype=torch.int64)}  # :552-555 (NO attention_mask)
        else:
            enc = self._tokenizer(prompts, padding=True, return_tensors="pt")  # :557-559
            inputs = {"input_ids": enc["input_ids"], "attention_mask": enc["attention_mask"]}
        S = inputs["input_ids"].shape[-1]
        num_prompt_tokens = S * inputs["input_ids"].shape[0]  # :560-562 (counts pads)
        if max_tokens is None:
            max_tokens = self.max_length - S  # :563-564
        if S + max_tokens > self.max_length:  # :565-572
            raise OracleError(
                f"This model's maximum context length is {self.max_length} tokens. "
                f"However, you requested {max_tokens + S} tokens "
                f"({S} in the messages, "
                f"{max_tokens} in the completion). "
                f"Please reduce the length of the messages or completion.")
        # :376-386
        if frequency_penalty is not None and frequency_penalty > 0:
            raise OracleError("'frequency_penalty' is not supported")
        if n is not None and n > 1:
            raise OracleError("'n' > 1 is not supported")
        generation_config = self.build_generation_config(max_tokens, top_p, temperature, presence_penalty)
        # :576-593 — `stop` defaults to [] in the vLLM request type, so a criterion is always attached
        crit = None
        stopping = None
        if stop is not None:
            stop_list = stop if isinstance(stop, (list, tuple)) else [stop]
            seqs = []
            for s in stop_list:
                if isinstance(s, str):
                    seqs.append(self._tokenizer.encode(s, return_tensors="pt", add_special_tokens=False)[0])
                else:  # token-id stop sequences (tokenizer-less synthetic runs)
                    seqs.append(torch.tensor(list(s), dtype=torch.int64))
            crit = StopSequenceStoppingCriteria(input_length=S, stop_sequences=seqs)
            stopping = StoppingCriteriaList([crit])

        # _handle_request :286-339 (non-stream branch)
        output_start = 0 if echo else S  # :324-327
        t0 = time.perf_counter()
        if want_logits:
            generation_config.return_dict_in_generate = True
            generation_config.output_logits = True
            generation_config.output_scores = True      # processed scores (== logits unless a processor is active)
        out = self._model.generate(**inputs, stopping_criteria=stopping, generation_config=generation_config)  # :328
        dt = time.perf_counter() - t0
        seqs_out = out.sequences if want_logits else out
        completion_tokens = seqs_out[:, S:].shape[-1] * seqs_out.shape[0]  # :329-335
        texts = None
        if self._tokenizer is not None:
            texts = self._tokenizer.batch_decode(seqs_out[:, output_start:], skip_special_tokens=True)  # :336-338
        finish_reason = "stop" if (crit is not None and crit.triggered) else "length"  # :621-627
        return OracleResult(output_ids=seqs_out, texts=texts, finish_reason=finish_reason,
                            prompt_tokens=num_prompt_tokens, completion_tokens=completion_tokens,
                            step_logits=[l.float() for l in out.logits] if want_logits else None,
                            seconds=dt,
                            extra={"step_scores": [x.float() for x in out.scores]} if want_logits else {})

    @torch.no_grad()
    def forward_logits(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Teacher-forced full forward; fp32 logits [B, S, V] (LlamaForCausalLM.forward, modeling_llama.py:445-499)."""
        kw = {}
        if attention_mask is not None:
            kw["attention_mask"] = attention_mask
            pos = attention_mask.long().cumsum(-1) - 1
            kw["position_ids"] = pos.masked_fill(attention_mask == 0, 0)
        return self._model(input_ids=input_ids, **kw).logits.float()
