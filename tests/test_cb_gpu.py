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


# ---------------------------------------------------------------------------------------------------------------------
# round 2: KV page pool, prefix reuse across requests, chunked prefill, per-sequence sampling, pool exhaustion
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def peaked():
    """tiny_g2 with the peaked LM head: every greedy step is decisive, so ids can be compared EXACTLY between the paths"""
    import torch
    c = load_case("tiny_g2_peaked")
    m = c["meta"]
    e = make_engine(m["cfg"], m["seed"], vocab_rows=m["vocab_rows"], max_batch=8, max_seq_len=512)
    g = torch.Generator().manual_seed(7)
    long_prompts = torch.randint(3, 1000, (3, 300), generator=g)
    long_prompts[1, :256] = long_prompts[0, :256]          # shares two 128-token blocks with prompt 0
    long_prompts[2, :128] = long_prompts[0, :128]          # shares one
    ref = e.generate(long_prompts, None, max_new_tokens=12, pad_token_id=m["pad_token_id"]).output_ids[:, 300:]
    yield e, c, long_prompts, ref
    e.close()


def test_prefix_reuse_skips_prefill_and_yields_identical_ids(peaked):
    """128-token blocks of a finished prefill stay in the pool (hash of the tokens up to the block ->
    reference-counted pages); a later prompt with the same leading blocks shares the pages, only its tail is computed,
    and its greedy ids are those of a full prefill."""
    e, c, prompts, ref = peaked
    rows = [r.tolist() for r in prompts]
    e.cb_begin(0, [])
    e.cb_config(0, True)
    try:
        s0 = e.cb_admit([rows[0]], [12])[0]
        st = e.cb_stats()
        assert st["prefix_hit_tokens"] == 0 and st["prefilled_tokens"] == 300 and st["cached_blocks"] == 2
        _run_until_done(e, [s0])
        assert e.cb_read(s0, 0, 12) == ref[0].tolist()
        # the same prompt again, while the first one still holds its slot: 256 of 300 tokens come from shared pages
        s1 = e.cb_admit([rows[0]], [12])[0]
        st = e.cb_stats()
        assert st["prefix_hit_tokens"] == 256 and st["prefilled_tokens"] == 300 + 44
        # ... and prompts that share two / one leading block(s), admitted together
        s2, s3 = e.cb_admit([rows[1], rows[2]], [12, 12])
        st = e.cb_stats()
        assert st["prefix_hit_tokens"] == 256 + 256 + 128 and st["prefilled_tokens"] == 300 + 44 + 44 + 172
        _run_until_done(e, [s1, s2, s3])
        assert e.cb_read(s1, 0, 12) == ref[0].tolist()
        assert e.cb_read(s2, 0, 12) == ref[1].tolist() and e.cb_read(s3, 0, 12) == ref[2].tolist()
        # releasing every sequence keeps the cached blocks alive (the cache holds its own references) ...
        for s in (s0, s1, s2, s3):
            e.cb_release(s)
        s4 = e.cb_admit([rows[0]], [12])[0]
        assert e.cb_stats()["prefix_hit_tokens"] == 256 + 256 + 128 + 256
        _run_until_done(e, [s4])
        assert e.cb_read(s4, 0, 12) == ref[0].tolist()
        e.cb_release(s4)
    finally:
        e.cb_end()


def test_chunked_prefill_interleaves_with_decode_and_matches_full_prefill(peaked):
    """with prefill chunks an admit is one chunk pass per decode step — the running sequence keeps
    generating while a long prompt is being prefilled — and the chunked prompt's ids equal those of a one-pass prefill."""
    e, c, prompts, ref = peaked
    rows = [r.tolist() for r in prompts]
    short = c["input_ids"][0].tolist()
    e.cb_begin(0, [])
    e.cb_config(128, False)
    try:
        a = e.cb_admit([short], [16])[0]
        e.cb_step(1)
        assert e.cb_poll()[0][a] == 2                           # one iteration = its only chunk (first token) + a decode step
        b = e.cb_admit([rows[0]], [12])[0]                       # 300 tokens = chunks of 128 + 128 + 44
        assert e.cb_stats()["pending_prompts"] == 1 and e.cb_poll()[0][b] == 0
        gen_a = []
        for _ in range(3):
            e.cb_step(1)
            n_gen, _, _ = e.cb_poll()
            gen_a.append(n_gen[a])
        assert gen_a == [3, 4, 5]                               # the short sequence never stalled
        assert e.cb_poll()[0][b] == 2 and e.cb_stats()["pending_prompts"] == 0   # first token from its last chunk + the same iteration's decode step
        _run_until_done(e, [a, b])
        assert e.cb_read(b, 0, 12) == ref[0].tolist()
        assert e.cb_read(a, 0, 16) == c["gen"][0].tolist()      # the oracle's ids (every step of this fixture is decisive)
        assert e.cb_stats()["prefill_passes"] == 4
    finally:
        e.cb_end()


def test_pool_exhaustion_admits_nothing_and_recovers():
    from kserve_b200.engine import PoolExhausted
    c = load_case("tiny_g2_peaked")
    m = c["meta"]
    e = make_engine(m["cfg"], m["seed"], vocab_rows=m["vocab_rows"], max_batch=4, max_seq_len=512, num_kv_pages=10)   # 8 pages = one full-length sequence
    try:
        prompts = [row.tolist() for row in c["input_ids"]]      # 48 tokens each
        e.cb_begin(0, [])
        a = e.cb_admit([prompts[0]], [200])[0]                   # 248 tokens -> 4 pages
        with pytest.raises(PoolExhausted):
            e.cb_admit([prompts[1], prompts[2]], [200, 200])     # 8 more pages: only 6 are left
        assert e.cb_stats()["available_pages"] == 6              # nothing leaked
        b = e.cb_admit([prompts[1]], [200])[0]
        e.cb_step(4)
        assert e.cb_read(a, 0, 5) == c["gen"][0][:5].tolist() and e.cb_read(b, 0, 5) == c["gen"][1][:5].tolist()
        e.cb_release(a)
        assert e.cb_stats()["available_pages"] == 6
        e.cb_end()
    finally:
        e.close()


def test_per_sequence_logits_processors_in_the_running_batch():
    """Sampling parameters ride along per sequence: a repetition-penalty row reproduces the oracle's greedy-with-penalty
    ids (fixture tiny_g2_reppen), next to a plain greedy row (fixture tiny_g2_ids) and a seeded sampling row."""
    cp, cg = load_case("tiny_g2_reppen"), load_case("tiny_g2_ids")
    m = cp["meta"]
    e = make_engine(m["cfg"], m["seed"], vocab_rows=m["vocab_rows"], max_batch=8, max_seq_len=512)
    try:
        pen = m["presence_penalty"]
        rows_p = [r.tolist() for r in cp["input_ids"]]
        e.cb_begin(0, [])
        outs = []
        for trial in range(2):
            slots = e.cb_admit([rows_p[0], cg["input_ids"][1].tolist(), rows_p[2], rows_p[3]], [cp["T"], cg["T"], cp["T"], 12],
                               sampling=[dict(repetition_penalty=pen), None, dict(repetition_penalty=pen),
                                         dict(do_sample=True, temperature=0.8, top_p=0.9, top_k=20, seed=1234)])
            _run_until_done(e, slots)
            outs.append([e.cb_read(s, 0, 64) for s in slots])
            for s in slots:
                e.cb_release(s)
        o = outs[0]
        tolp, tolg = logits_tol(cp["step_logits"]), logits_tol(cg["step_logits"])
        for row, fix_row in ((0, 0), (2, 2)):                  # penalty rows == the oracle's, up to a non-decisive step
            ref = cp["gen"][fix_row].tolist()
            for t, (x, y) in enumerate(zip(o[row], ref)):
                if x != y:
                    assert float(cp["margin"][fix_row, t]) <= 2 * tolp, f"penalty row {row} diverges at decisive step {t}"
                    break
        ref = cg["gen"][1].tolist()
        for t, (x, y) in enumerate(zip(o[1], ref)):
            if x != y:
                assert float(cg["margin"][1, t]) <= 2 * tolg
                break
        assert len(o[3]) == 12 and outs[1][3] == o[3]           # the seed fixes the sampled stream, whoever shares the batch
        assert outs[1][:3] == o[:3]
        e.cb_end()
    finally:
        e.close()


def test_host_dram_kv_tier_preempt_and_resume_bit_identical(peaked):
    """BASELINE configs[3] "paged-KV with host-DRAM KV-offload tier" in the continuous batcher: a running sequence is
    swapped out to pinned host memory (its pages go back to the pool and are overwritten by other sequences), swapped back
    in on fresh pages, and continues with exactly the ids of an uninterrupted run."""
    e, c, prompts, ref = peaked
    rows = [r.tolist() for r in prompts]
    short = [r.tolist() for r in c["input_ids"]]
    e.cb_begin(0, [])
    try:
        a = e.cb_admit([rows[0]], [12])[0]
        e.cb_step(3)                                           # 4 tokens generated
        free0 = e.cb_stats()["available_pages"]
        e.cb_swap_out(a)
        st = e.cb_stats()
        assert st["swap_outs"] == 1 and st["available_pages"] == free0 + 5      # 300 + 12 tokens -> 5 pages back in the pool
        others = e.cb_admit(short[:3], [16, 16, 16])            # reuse (and overwrite) the freed pages
        e.cb_step(6)
        assert e.cb_poll()[0][a] == 4                           # the swapped sequence does not advance
        e.cb_swap_in(a)
        assert e.cb_stats()["swap_ins"] == 1
        _run_until_done(e, [a] + others)
        assert e.cb_read(a, 0, 12) == ref[0].tolist()
        for i, s in enumerate(others):
            assert e.cb_read(s, 0, 16) == c["gen"][i].tolist()
    finally:
        e.cb_end()


def test_scheduler_preempts_to_host_when_the_pool_is_exhausted():
    from kserve_b200.continuous import ContinuousBatcher
    c = load_case("tiny_g2_peaked")
    m = c["meta"]
    e = make_engine(m["cfg"], m["seed"], vocab_rows=m["vocab_rows"], max_batch=8, max_seq_len=512, num_kv_pages=16)
    try:
        ids = c["input_ids"]
        prompts = [r.tolist() for r in ids]
        cb = ContinuousBatcher(e, pad_token_id=0, eos_token_ids=[], steps_per_poll=2, kv_offload=True)
        cb.start()

        async def main():          # each request needs 7 pages (48 + 400 tokens): the third admission exhausts the 16-page pool
            return await asyncio.gather(*[cb.submit([prompts[i]], ids[i:i + 1], 400) for i in range(3)])
        try:
            rs = asyncio.run(main())
        finally:
            cb.stop()
        for i, r in enumerate(rs):
            assert r.num_generated == 400
            assert r.output_ids[0, ids.shape[1]:ids.shape[1] + c["T"]].tolist() == c["gen"][i].tolist()
        assert cb.stats.get("preempted", 0) >= 1 and cb.stats.get("resumed", 0) == cb.stats["preempted"]
    finally:
        e.close()


def test_boundary_lengths_through_the_continuous_batcher_match_the_oracle():
    """the boundary-length oracle fixture (rows of 1 / 63 / 64 / 65 / 127 / 128 prompt tokens, 64 generated each) through
    b200_cb_*: two rows start, four join the running batch, prompts go through 128-token chunks with the prefix cache on,
    the KV pool holds exactly the 15 pages these six sequences need and max_seq_len is the longest row's 192 tokens.
    Every step of the fixture is decisive, so each sequence's ids must equal the oracle's token for token."""
    c = load_case("tiny_g4_peaked_edges")
    m = c["meta"]
    T, B = c["T"], c["input_ids"].shape[0]
    rows = [c["input_ids"][b][c["mask"][b].bool()].tolist() for b in range(B)]
    assert [len(r) for r in rows] == [1, 63, 64, 65, 127, 128] and T == 64
    pages = sum(-(-(len(r) + T) // 64) for r in rows)
    assert pages == 15
    e = make_engine(m["cfg"], m["seed"], vocab_rows=m["vocab_rows"], max_batch=B, max_seq_len=192, num_kv_pages=pages)
    try:
        e.cb_begin(m["pad_token_id"], [])
        e.cb_config(128, True)
        slots = e.cb_admit(rows[:2], [T, T])
        e.cb_step(5)
        slots += e.cb_admit(rows[2:], [T] * 4)                # join a running batch; the pool is now fully committed
        assert e.cb_stats()["available_pages"] == 0
        n_gen, fin, stop = _run_until_done(e, slots)
        for b, s in enumerate(slots):
            assert n_gen[s] == T and not stop[s]
            assert e.cb_read(s, 0, T) == c["gen"][b].tolist(), f"row {b} ({len(rows[b])} prompt tokens)"
        for s in slots:
            e.cb_release(s)
        assert e.cb_stats()["available_pages"] == pages      # nothing leaked (the cached 128-token block is evictable)
    finally:
        e.cb_end()
        e.close()
