"""Shared helpers for the parity tests (golden fixture loading, engine construction)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated bf16 tolerance for logits: both sides run a bf16 pipeline with fp32
# accumulation; they differ in accumulation order only.  One bf16 ulp is 2^-8 relative; after L layers of
# residual updates the logits of the two pipelines are observed to differ by a few ulps of the largest
# logit.  TOL_ULPS * 2^-8 * max|logit| is the absolute tolerance used everywhere below.
TOL_ULPS = 4.0


def logits_tol(ref_logits: torch.Tensor) -> float:
    return TOL_ULPS * 2.0 ** -8 * float(ref_logits.abs().max())


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    out = torch.from_numpy(z["output_ids"].astype(np.int64))
    T = meta["completion_tokens"] // out.shape[0]
    S = out.shape[1] - T
    case = dict(meta=meta, output_ids=out, S=S, T=T, input_ids=out[:, :S].contiguous(), gen=out[:, S:].contiguous(),
                topk_vals=torch.from_numpy(z["topk_vals"]), topk_idx=torch.from_numpy(z["topk_idx"]))
    if "step_logits" in z:
        case["step_logits"] = torch.from_numpy(z["step_logits"])  # [B, T, V]
    if "step0_logits" in z:
        case["step0_logits"] = torch.from_numpy(z["step0_logits"].astype(np.float32))
    pad = meta["pad_token_id"]
    mask = (case["input_ids"] != pad).long() if pad is not None else torch.ones_like(case["input_ids"])
    case["mask"] = mask
    case["margin"] = case["topk_vals"][..., 0] - case["topk_vals"][..., 1]  # [B, T]
    return case


def make_engine(cfg_name, seed, vocab_rows=None, **kw):
    from oracle import weights as W  # the oracle package also owns the synthetic checkpoint generator
    from kserve_b200.engine import B200Engine
    cfg = W.CONFIGS[cfg_name]
    eng = B200Engine(cfg, vocab_rows=vocab_rows, **kw)

    def gen():
        for name, t in W.iter_state_dict(cfg, seed):
            if vocab_rows and name in ("model.embed_tokens.weight", "lm_head.weight"):
                t = t[:vocab_rows] if t.shape[0] >= vocab_rows else torch.cat(
                    [t, torch.zeros(vocab_rows - t.shape[0], t.shape[1], dtype=t.dtype)])
            yield name, t
    eng.load_weights(gen())
    return eng
