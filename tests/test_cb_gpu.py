"""Continuous (iteration-level) batching on a real GPU — b200_cb_* through the C ABI and the Python scheduler.

Parity rule: a sequence that joins a running batch must produce the oracle's greedy tokens for that prompt (same
margin-aware comparison as test_engine_gpu.py: ids equal up to the first step whose oracle margin is inside the
stated bf16 tolerance), whatever else is in the batch and whenever it joined.
"""
import asyncio
import os

import pytest

from helpers import GOLDEN, load_case, logits_tol, make_engine

pytestmark = pytest.mark.gpu


def _check_row(name, c, b, toks, upto=None):
    tol = logits_tol(c["step_logits"])
    ref = c["gen"][b].tolist()
    n = len(toks) if upto is None else upto
    for t in range(min(n, len(toks))):
        if toks[t] != ref[t]:
            assert float(c["margin"][b, t]) <= 2 * tol, f"{name} row {b} diverges at decisive step {t}"
            return t
    return min(n, len(toks))


@pytest.fixture(scope="module")
def eng():
    c = load_case("tiny_g2_ids")
    m = c["meta"]
    e = make_engine(m["cfg"], m["seed"], vocab_rows=m["vocab_rows"], max_batch=8, max_seq_len=512)
    yield e, c
    e.close()


def _run_until_done(e, slots, max_iters=400):
    for _ in range(max_iters):
        n_gen, fin, stop = e.cb_poll()
        if all(fin[s] for s in slots):
            return n_gen, fin, stop
        e.cb_step(1)
    raise AssertionError("sequences did not finish")


def test_join_running_batch_matches_oracle(eng):
    e, c = eng
    T, B = c["T"], c["input_ids"].shape[0]
    prompts = [row.tolist() for row in c["input_ids"]]
    e.cb_begin(c["meta"]["pad_token_id"] or 0, [])
    try:
        s0 = e.cb_admit([prompts[0]], [T])
        e.cb_step(3)                                   # row 0 is 4 tokens in when the others join
        n_gen, fin, _ = e.cb_poll()
        assert n_gen[s0[0]] == 4 and not fin[s0[0]]
        rest = e.cb_admit(prompts[1:], [T] * (B - 1))
        slots = s0 + rest
        assert len(set(slots)) == B
        n_gen, fin, stop = _run_until_done(e, slots)
        for b, s in enumerate(slots):
            assert n_gen[s] == T and not stop[s]
            _check_row("join", c, b, e.cb_read(s, 0, T))
        # a static-batch generate must be refused while the mode is on, and work again after cb_end
        with pytest.raises(Exception):
            e.generate(c["input_ids"], None, max_new_tokens=2, pad_token_id=0)
        for s in slots:
            e.cb_release(s)
    finally:
        e.cb_end()
    r = e.generate(c["input_ids"], None, max_new_tokens=T, pad_token_id=c["meta"]["pad_token_id"])
    assert r.num_generated == T


def test_eos_stop_sequences_and_slot_reuse(eng):
    e, c = eng
    T = c["T"]
    prompts = [row.tolist() for row in c["input_ids"]]
    g0, g1 = c["gen"][0].tolist(), c["gen"][1].tolist()
    eos = g0[3]
    n_eos = g0.index(eos) + 1                          # the EOS token itself is emitted (utils.py:2796-2797)
    stop_seq = g1[2:4]
    # first step at which the generated suffix of row 1 equals the stop sequence
    n_stop = next(k + 2 for k in range(len(g1) - 1) if g1[k:k + 2] == stop_seq)
    e.cb_begin(0, [eos])
    try:
        a = e.cb_admit([prompts[0]], [T])[0]
        b = e.cb_admit([prompts[1]], [T], [[stop_seq]])[0]
        n_gen, fin, stop = _run_until_done(e, [a, b])
        toks_a, toks_b = e.cb_read(a, 0, T), e.cb_read(b, 0, T)
        if _check_row("eos", c, 0, toks_a, upto=n_eos) == n_eos:
            assert n_gen[a] == n_eos and toks_a[-1] == eos and not stop[a]
        n_b = min(n_stop, (g1.index(eos) + 1) if eos in g1 else T)
        if _check_row("stop", c, 1, toks_b, upto=n_b) == n_b:
            assert n_gen[b] == n_b
            assert bool(stop[b]) == (n_b == n_stop)
        # length limit, in a slot that was used before
        e.cb_release(a)
        a2 = e.cb_admit([prompts[2]], [5])[0]
        assert a2 == a
        n_gen, fin, stop = _run_until_done(e, [a2])
        g2 = c["gen"][2].tolist()
        n2 = min(5, (g2.index(eos) + 1) if eos in g2[:5] else 5)
        if _check_row("reuse", c, 2, e.cb_read(a2, 0, 5), upto=n2) == n2:
            assert n_gen[a2] == n2
    finally:
        e.cb_end()


def test_scheduler_concurrent_requests(eng):
    """ContinuousBatcher: three requests in flight at once (one of them a two-row request with a batch-wide stop)."""
    from kserve_b200.continuous import ContinuousBatcher
    e, c = eng
    T = c["T"]
    ids = c["input_ids"]
    prompts = [row.tolist() for row in ids]
    g1 = c["gen"][1].tolist()
    stop_seq = g1[4:6]
    n_stop = next(k + 2 for k in range(len(g1) - 1) if g1[k:k + 2] == stop_seq)
    cb = ContinuousBatcher(e, pad_token_id=0, eos_token_ids=[], steps_per_poll=3)
    cb.start()
    streamed = []

    async def main():
        t1 = asyncio.create_task(cb.submit([prompts[0]], ids[0:1], T))
        await asyncio.sleep(0.05)
        t2 = asyncio.create_task(cb.submit(prompts[1:3], ids[1:3], T, [stop_seq]))
        t3 = asyncio.create_task(cb.submit([prompts[3 % len(prompts)]], ids[3 % len(prompts)][None], 6,
                                           on_tokens=lambda step, toks: streamed.append((step, toks[0]))))
        return await asyncio.gather(t1, t2, t3)
    try:
        r1, r2, r3 = asyncio.run(main())
    finally:
        cb.stop()
    S = ids.shape[1]
    assert r1.num_generated == T and not r1.stop_triggered and r1.output_ids.shape == (1, S + T)
    _check_row("sched r1", c, 0, r1.output_ids[0, S:].tolist())
    # the stop matched in row 1: the whole request (rows 1 and 2) ends at that step
    k = _check_row("sched r2 row1", c, 1, r2.output_ids[0, S:].tolist(), upto=min(n_stop, r2.num_generated))
    if k == n_stop:
        assert r2.stop_triggered and r2.num_generated == n_stop and r2.output_ids.shape == (2, S + n_stop)
        _check_row("sched r2 row2", c, 2, r2.output_ids[1, S:].tolist(), upto=n_stop)
    b3 = 3 % len(prompts)
    assert r3.num_generated == 6
    _check_row("sched r3", c, b3, r3.output_ids[0, S:].tolist())
    assert [s for s, _ in streamed] == list(range(6)) and [t for _, t in streamed] == r3.output_ids[0, S:].tolist()
    assert cb.stats["finished"] == 3 and cb.free_slots == e.max_batch


def test_openai_route_with_continuous_batching():
    """The OpenAI completions route served by the continuous batcher returns the oracle's strings and usage."""
    from fastapi.testclient import TestClient
    from transformers import AutoTokenizer
    from kserve_b200.generative_model import B200GenerativeModel
    from kserve_b200.kserve_api import ModelServer
    from oracle import weights as W
    c = load_case("tiny_g2_text")
    m = c["meta"]
    cfg = dict(W.CONFIGS["tiny_g2"], architectures=["LlamaForCausalLM"], model_type="llama")
    tok = AutoTokenizer.from_pretrained(os.path.join(GOLDEN, "byte_tokenizer"))
    model = B200GenerativeModel("tiny", model_config=cfg, state_dict=W.iter_state_dict(W.CONFIGS["tiny_g2"], 0), tokenizer=tok,
                                max_model_len=512, max_batch=8, continuous_batching=True)
    assert model.load() and model._cb is not None
    try:
        with TestClient(ModelServer().create_application([model])) as client:
            r = client.post("/openai/v1/completions", json={"model": "tiny", "prompt": m["prompt"], "max_tokens": m["max_tokens"], "temperature": 0})
            assert r.status_code == 200, r.text
            j = r.json()
            assert [ch["text"] for ch in j["choices"]] == m["texts"]
            assert j["usage"] == {"prompt_tokens": m["prompt_tokens"], "completion_tokens": m["completion_tokens"],
                                  "total_tokens": m["prompt_tokens"] + m["completion_tokens"]}
            with client.stream("POST", "/openai/v1/completions", json={"model": "tiny", "prompt": m["prompt"][1], "max_tokens": 8, "stream": True}) as s:
                body = "".join(s.iter_text())
            assert body.strip().endswith("data: [DONE]")
            pred = client.post("/v1/models/tiny:predict", json={"instances": [c["input_ids"][1].tolist()], "parameters": {"max_tokens": 4}})
            assert pred.status_code == 200 and len(pred.json()["predictions"][0]) == 4
    finally:
        model.stop()
