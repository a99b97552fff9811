"""Generate tests/golden/* by running the oracle (transformers CPU backend) here.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Run from the repo root:

    python -m oracle.make_goldens            # all small cases (~1 min)
    python -m oracle.make_goldens --big      # + llama3_8b_2l case (~few min, 10 GB RAM)

The fixtures are committed; the GPU tests compare the CUDA engine with them and
the CPU tests re-run the oracle against them (determinism of the checker).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import transformers

from . import weights as W
from .hf_oracle import OracleGenerativeModel, build_llama

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CHAT_TEMPLATE = (
    "{% for message in messages %}<|{{ message['role'] }}|>{{ message['content'] }}\n{% endfor %}"
    "{% if add_generation_prompt %}<|assistant|>{% endif %}"
)


def build_byte_tokenizer(out_dir: str):
    """A 256-byte + specials byte-level tokenizer built from scratch (no Hub).

    It has NO pad token on purpose, so loading it exercises the reference's
    fallback ``[PAD]`` + ``resize_token_embeddings`` path
    (generative_model.py:256-265, tests/test_model.py:502-509).
    """
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast

    alphabet = sorted(pre_tokenizers.ByteLevel.alphabet())
    vocab = {ch: i for i, ch in enumerate(alphabet)}
    tok = Tokenizer(models.BPE(vocab=vocab, merges=[]))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)
    tok.decoder = decoders.ByteLevel()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, bos_token=None, eos_token=None, unk_token=None)
    fast.chat_template = CHAT_TEMPLATE
    os.makedirs(out_dir, exist_ok=True)
    fast.save_pretrained(out_dir)
    return fast


def load_byte_tokenizer():
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(os.path.join(GOLDEN, "byte_tokenizer"))


def topk_pack(step_logits, k=16):
    vals, idx = [], []
    for l in step_logits:
        v, i = torch.topk(l, k, dim=-1)
        vals.append(v.numpy())
        idx.append(i.numpy())
    return np.stack(vals, 1), np.stack(idx, 1)  # [B, T, k]


def gen_case(name, cfg_name, seed, prompt, max_tokens, stop=(), tokenizer=None, pad_token_id=None,
             full_logits=True, echo=False, stop_from=None):
    cfg = W.CONFIGS[cfg_name]
    sd = W.synth_state_dict(cfg, seed)
    model = build_llama(cfg, sd)
    orc = OracleGenerativeModel(model, tokenizer=tokenizer, pad_token_id=pad_token_id)
    if stop_from is not None:  # derive a stop sequence that is guaranteed to occur: (row, i, j) of a free run
        row, i, j = stop_from
        free = orc.create_completion(prompt, max_tokens=max_tokens, temperature=0)
        S = free.output_ids.shape[1] - max_tokens
        stop = [free.output_ids[row, S + i:S + j].tolist()]
    res = orc.create_completion(prompt, max_tokens=max_tokens, stop=stop, echo=echo, temperature=0,
                                want_logits=True)
    tv, ti = topk_pack(res.step_logits)
    out = dict(
        output_ids=res.output_ids.numpy(),
        topk_vals=tv, topk_idx=ti,
        meta=json.dumps(dict(
            name=name, cfg=cfg_name, seed=seed, max_tokens=max_tokens,
            finish_reason=res.finish_reason, prompt_tokens=res.prompt_tokens,
            completion_tokens=res.completion_tokens, texts=res.texts,
            pad_token_id=orc.pad_token_id, vocab_rows=int(model.get_input_embeddings().weight.shape[0]),
            weights_checksum=W.checksum(sd), stop=[list(s) if not isinstance(s, str) else s for s in stop],
            prompt=prompt if isinstance(prompt, (str, list)) and (isinstance(prompt, str) or isinstance(prompt[0], str)) else None,
            echo=echo,
            transformers=transformers.__version__, torch=torch.__version__,
            threads=torch.get_num_threads())),
    )
    if full_logits:
        out["step_logits"] = torch.stack(res.step_logits, 1).numpy().astype(np.float32)  # [B, T, V]
    else:
        out["step0_logits"] = res.step_logits[0].numpy().astype(np.float16)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: ids {out['output_ids'].shape} finish={res.finish_reason} "
          f"{res.seconds:.2f}s  margins(top1-top2) min={float((tv[..., 0]-tv[..., 1]).min()):.4f}")


def ids_prompt(B, S, vocab_hi, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(3, vocab_hi, (B, S), generator=g).tolist()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    build_byte_tokenizer(os.path.join(GOLDEN, "byte_tokenizer"))

    # (a) equal-length token-id prompts (generative_model.py:552-555 path), group size 2
    gen_case("tiny_g2_ids", "tiny_g2", 0, ids_prompt(4, 48, 1000, 1234), 16, pad_token_id=1030)
    # (b) ragged string prompts -> tokenizer, left padding, fallback [PAD] + embedding resize
    tok = load_byte_tokenizer()
    gen_case("tiny_g2_text", "tiny_g2", 0,
             ["Hello world", "The quick brown fox jumps over the lazy dog.", "KServe on B200!", "a"],
             12, tokenizer=tok)
    # (c) stop sequence as token ids: take case (a)'s 5th+6th generated token of row 2 -> must stop the batch
    gen_case("tiny_g2_stop", "tiny_g2", 0, ids_prompt(4, 48, 1000, 1234), 16, stop_from=(2, 4, 6), pad_token_id=1030)
    # (d) group size 4, 3 layers, odd batch, longer prompt (crosses KV page boundary of 64)
    gen_case("tiny_g4_ids", "tiny_g4", 1, ids_prompt(3, 100, 2048, 99), 40, pad_token_id=0)
    # (e) single sequence, batch 1, long decode crossing pages
    gen_case("tiny_g4_b1", "tiny_g4", 1, ids_prompt(1, 17, 2048, 5), 80, pad_token_id=0)
    # (f) token-id prompts that contain the pad id => transformers infers a mask (quirk q4)
    p = ids_prompt(2, 32, 1000, 77)
    p[0][:5] = [1030] * 5
    gen_case("tiny_g2_padinfer", "tiny_g2", 0, p, 8, pad_token_id=1030)
    if args.big:
        gen_case("llama3_8b_2l_ids", "llama3_8b_2l", 0, ids_prompt(2, 96, 128000, 1234), 8,
                 pad_token_id=128255, full_logits=False)


if __name__ == "__main__":
    sys.exit(main())
