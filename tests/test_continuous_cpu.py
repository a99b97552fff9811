"""Host logic of the continuous batcher against a scripted engine (no GPU): admission under slot pressure,
lockstep rows, batch-wide stop truncation, EOS padding, streaming order, slot accounting."""
import asyncio

import torch

from kserve_b200.continuous import ContinuousBatcher


class ScriptedEngine:
    """cb_* surface of B200Engine; sequence i emits script[i][k] as its k-th token."""
    def __init__(self, max_batch, scripts, eos=(), step_delay=0.0, pages=None, first_token_after=0):
        self.step_delay = step_delay
        self.pages = pages                # KV pool size in "pages" (1 page per prompt token here); None = unlimited
        self.first_token_after = first_token_after   # chunked prefill: the first token appears after this many steps
        self.configured = None
        self.max_batch, self.max_seq_len = max_batch, 4096
        self.scripts = scripts            # prompt tuple -> token list
        self.eos = set(eos)
        self.slots = {}
        self.calls = []

    def cb_begin(self, pad, eos):
        self.eos = set(eos)

    def cb_end(self):
        pass

    def _emit(self, st):
        if st["fin"]:
            return
        tok = st["script"][len(st["out"])]
        st["out"].append(tok)
        if tok in self.eos or len(st["out"]) >= st["max_new"]:
            st["fin"] = True
        for q in st["stops"]:
            if len(q) <= len(st["out"]) and st["out"][-len(q):] == q:
                st["fin"] = st["stop"] = True

    def cb_config(self, chunk, prefix):
        self.configured = (chunk, prefix)

    def cb_admit(self, prompts, max_new, stops, sampling=None):
        from kserve_b200.engine import PoolExhausted
        free = [s for s in range(self.max_batch) if s not in self.slots]
        assert len(prompts) <= len(free)
        if self.pages is not None:
            used = sum(st["pages"] for st in self.slots.values())
            if used + sum(len(p) for p in prompts) > self.pages:
                self.calls.append(("exhausted", len(prompts)))
                raise PoolExhausted("KV page pool exhausted")
        out = []
        for i, (p, m, ss) in enumerate(zip(prompts, max_new, stops)):
            s = free.pop(0)
            self.slots[s] = dict(script=self.scripts[tuple(p)], out=[], max_new=m, stops=[list(q) for q in ss], fin=False, stop=False,
                                 pages=len(p), wait=self.first_token_after, sampling=(sampling or [None] * len(prompts))[i])
            if not self.first_token_after:
                self._emit(self.slots[s])
            out.append(s)
        self.calls.append(("admit", len(prompts)))
        return out

    def cb_step(self, n):
        if self.step_delay:
            import time
            time.sleep(self.step_delay * n)
        for _ in range(n):
            for st in self.slots.values():
                if st.get("swapped"):
                    continue                  # its KV is in host DRAM: not part of the decode batch
                if st["wait"] > 0:
                    st["wait"] -= 1           # still being prefilled (one chunk per step)
                    if st["wait"] > 0:
                        continue
                self._emit(st)
        self.calls.append(("step", n, len(self.slots)))

    def cb_poll(self):
        g = [len(self.slots[s]["out"]) if s in self.slots else 0 for s in range(self.max_batch)]
        f = [int(self.slots[s]["fin"]) if s in self.slots else 0 for s in range(self.max_batch)]
        h = [int(self.slots[s]["stop"]) if s in self.slots else 0 for s in range(self.max_batch)]
        return g, f, h

    def cb_read(self, slot, first=0, cap=4096):
        return self.slots[slot]["out"][first:first + cap]

    def cb_release(self, slot):
        del self.slots[slot]

    def cb_swap_out(self, slot):
        st = self.slots[slot]
        assert not st.get("swapped")
        st["swapped"], st["host_pages"], st["pages"] = True, st["pages"], 0
        self.calls.append(("swap_out", slot))

    def cb_swap_in(self, slot):
        from kserve_b200.engine import PoolExhausted
        st = self.slots[slot]
        if self.pages is not None and sum(x["pages"] for x in self.slots.values()) + st["host_pages"] > self.pages:
            raise PoolExhausted("KV page pool exhausted")
        st["swapped"], st["pages"] = False, st["host_pages"]
        self.calls.append(("swap_in", slot))


def _run(cb, coro):
    cb.start()
    try:
        return asyncio.run(coro)
    finally:
        cb.stop()


def test_requests_queue_for_slots_and_all_complete():
    scripts = {(i,): [100 * i + k for k in range(40)] for i in range(6)}
    eng = ScriptedEngine(2, scripts)
    cb = ContinuousBatcher(eng, pad_token_id=0, steps_per_poll=3)

    async def main():
        return await asyncio.gather(*[cb.submit([[i]], torch.tensor([[i]]), 5 + i) for i in range(6)])
    res = _run(cb, main())
    for i, r in enumerate(res):
        assert r.num_generated == 5 + i and not r.stop_triggered
        assert r.output_ids.tolist() == [[i] + scripts[(i,)][:5 + i]]
    assert cb.free_slots == 2 and not eng.slots and cb.stats["finished"] == 6
    assert max(c[2] for c in eng.calls if c[0] == "step") <= 2          # never more rows than slots


def test_batchwide_stop_truncates_every_row_and_eos_pads():
    scripts = {(1,): [5, 6, 7, 8, 9, 10, 11, 12], (2,): [20, 21, 22, 23, 24, 25, 26, 27], (3,): [30, 99, 31, 32, 33, 34, 35, 36],
               (4,): [40, 41, 42, 43, 44, 45, 46, 47]}
    eng = ScriptedEngine(8, scripts)
    cb = ContinuousBatcher(eng, pad_token_id=0, eos_token_ids=[99], steps_per_poll=4)
    seen = []

    async def main():
        a = cb.submit([[1], [2]], torch.tensor([[1], [2]]), 8, [[7, 8]])             # stop matches in row 0 at step 4
        b = cb.submit([[3], [4]], torch.tensor([[3], [4]]), 6, on_tokens=lambda s, t: seen.append((s, list(t))))   # row 0 hits EOS at step 2
        return await asyncio.gather(a, b)
    ra, rb = _run(cb, main())
    assert ra.stop_triggered and ra.num_generated == 4
    assert ra.output_ids.tolist() == [[1, 5, 6, 7, 8], [2, 20, 21, 22, 23]]
    assert not rb.stop_triggered and rb.num_generated == 6
    assert rb.output_ids.tolist() == [[3, 30, 99, 0, 0, 0, 0], [4, 40, 41, 42, 43, 44, 45]]
    assert [s for s, _ in seen] == list(range(6))
    assert [t for _, t in seen] == [[30, 40], [99, 41], [0, 42], [0, 43], [0, 44], [0, 45]]


def test_invalid_requests_are_rejected_up_front():
    eng = ScriptedEngine(2, {})
    cb = ContinuousBatcher(eng)
    for bad in (dict(prompts=[[1]] * 3, padded=torch.zeros(3, 1), max_new_tokens=4, stop_sequences=[]),
                dict(prompts=[[1]], padded=torch.zeros(1, 1), max_new_tokens=5000, stop_sequences=[]),
                dict(prompts=[[1]], padded=torch.zeros(1, 1), max_new_tokens=4, stop_sequences=[[1] * 9])):
        try:
            cb.submit_nowait(done=lambda r, e: None, **bad)
            raise AssertionError("accepted an invalid request")
        except ValueError:
            pass


def test_cancelled_requests_leave_the_batch_and_free_their_slots():
    scripts = {(i,): list(range(1000 * i, 1000 * i + 400)) for i in range(3)}
    eng = ScriptedEngine(2, scripts, step_delay=0.002)
    cb = ContinuousBatcher(eng, pad_token_id=0, steps_per_poll=2)

    async def main():
        long_a = asyncio.create_task(cb.submit([[0]], torch.tensor([[0]]), 300))
        long_b = asyncio.create_task(cb.submit([[1]], torch.tensor([[1]]), 300))
        await asyncio.sleep(0.05)                         # both are running, the third request has to queue
        waiting = asyncio.create_task(cb.submit([[2]], torch.tensor([[2]]), 4))
        await asyncio.sleep(0.02)
        long_a.cancel()                                   # client went away mid-generation
        r = await waiting                                 # ... which frees a slot for the queued request
        long_b.cancel()
        for t in (long_a, long_b):
            try:
                await t
            except asyncio.CancelledError:
                pass
        return r
    r = _run(cb, main())
    assert r.output_ids.tolist() == [[2, 2000, 2001, 2002, 2003]]
    assert cb.stats["cancelled"] >= 1


def test_randomised_arrivals_keep_every_invariant():
    """200 requests with random sizes, lengths, EOS positions, stop sequences and a few cancellations: every finished
    request returns exactly the prefix of its script that the reference semantics dictate, no slot leaks."""
    import random
    rng = random.Random(7)
    EOS = 9999
    scripts, specs = {}, []
    for i in range(200):
        rows = rng.choice([1, 1, 1, 2, 3])
        max_new = rng.randint(1, 40)
        prompts = []
        for r in range(rows):
            key = (i, r)
            sc = [rng.randint(10, 500) for _ in range(64)]
            if rng.random() < 0.3:
                sc[rng.randint(0, 45)] = EOS
            scripts[key] = sc
            prompts.append(list(key))
        stop = []
        if rng.random() < 0.25:
            k = rng.randint(0, 30)
            stop = [scripts[(i, 0)][k:k + 2]]
        specs.append((prompts, max_new, stop, rng.random() < 0.05))
    eng = ScriptedEngine(8, scripts, step_delay=0.0002)
    cb = ContinuousBatcher(eng, pad_token_id=0, eos_token_ids=[EOS], steps_per_poll=3)

    def expected(prompts, max_new, stop):
        outs, stop_at = [], None
        for p in prompts:
            sc, o = scripts[tuple(p)], []
            for t in sc[:max_new]:
                o.append(t)
                if t == EOS:
                    break
                if stop and len(o) >= 2 and o[-2:] == stop[0]:
                    stop_at = len(o) if stop_at is None else min(stop_at, len(o))
                    break
            outs.append(o)
        if stop_at is not None:
            n = stop_at
            outs = [o[:n] for o in outs]
        else:
            n = max(len(o) for o in outs)
        return [o + [0] * (n - len(o)) for o in outs], stop_at is not None

    async def one(spec):
        prompts, max_new, stop, cancel = spec
        t = asyncio.create_task(cb.submit(prompts, torch.tensor(prompts), max_new, stop))
        if cancel:
            await asyncio.sleep(0.001)
            t.cancel()
        try:
            return await t
        except asyncio.CancelledError:
            return None

    async def main():
        return await asyncio.gather(*[one(s) for s in specs])
    results = _run(cb, main())
    done = 0
    for spec, r in zip(specs, results):
        if r is None:
            continue
        want, stopped = expected(*spec[:3])
        assert r.output_ids[:, 2:].tolist() == want and r.stop_triggered == stopped and r.num_generated == len(want[0])
        done += 1
    assert done >= 180 and cb.free_slots == 8 and not eng.slots
    assert cb.stats["finished"] + cb.stats["cancelled"] >= done


def test_pool_exhaustion_defers_admission_until_a_release():
    """b200_cb_admit answers "KV page pool exhausted" (nothing admitted): the scheduler keeps the request queued and
    retries after the next release instead of failing it; a request that can never fit an idle pool is failed."""
    scripts = {(1,) * 6: list(range(10, 20)), (2,) * 6: list(range(20, 30)), (3,) * 50: list(range(30, 40))}
    eng = ScriptedEngine(4, scripts, pages=10)       # two 6-token prompts do not fit together
    cb = ContinuousBatcher(eng, steps_per_poll=1, kv_offload=False)

    async def main():
        pad = lambda p: torch.tensor([p])
        a = cb.submit([[1] * 6], pad([1] * 6), 5)
        b = cb.submit([[2] * 6], pad([2] * 6), 5)
        ra, rb = await asyncio.gather(a, b)
        assert ra.output_ids[0, 6:].tolist() == list(range(10, 15)) and rb.output_ids[0, 6:].tolist() == list(range(20, 25))
        try:
            await cb.submit([[3] * 50], pad([3] * 50), 5)
        except ValueError as e:
            return str(e)
    msg = _run(cb, main())
    assert "does not fit the KV page pool" in msg
    kinds = [c[0] for c in eng.calls]
    assert "exhausted" in kinds and kinds.count("admit") == 2       # the second request waited for the first one's release


def test_chunked_prefill_config_sampling_passthrough_and_first_token_time():
    scripts = {(7, 8, 9): list(range(50, 60))}
    eng = ScriptedEngine(2, scripts, first_token_after=3)
    cb = ContinuousBatcher(eng, steps_per_poll=1, prefill_chunk_tokens=256, prefix_cache=True)

    async def main():
        return await cb.submit([[7, 8, 9]], torch.tensor([[7, 8, 9]]), 4, sampling=dict(do_sample=True, temperature=0.7, seed=5))
    r = _run(cb, main())
    assert eng.configured == (256, True)
    assert r.output_ids[0, 3:].tolist() == [50, 51, 52, 53] and r.num_generated == 4
    assert r.prefill_ms > 0 and r.decode_ms >= 0       # first token observed some steps after the admit


def test_replicated_engine_broadcasts_every_scheduler_call(monkeypatch):
    """Tensor parallel: the scheduler's b200_cb_* calls go to the follower ranks first (same order, same arguments)."""
    from kserve_b200 import continuous, tp
    sent = []
    monkeypatch.setattr(tp, "leader_call", lambda m, a, k: sent.append((m, a, k)))
    eng = ScriptedEngine(2, {(1, 2): [5, 6, 7]})
    eng.tp_size = 2
    cb = ContinuousBatcher(eng, steps_per_poll=1)
    assert isinstance(cb.engine, continuous.ReplicatedEngine)

    async def main():
        return await cb.submit([[1, 2]], torch.tensor([[1, 2]]), 3)
    r = _run(cb, main())
    assert r.output_ids[0, 2:].tolist() == [5, 6, 7]
    names = [m for m, _, _ in sent]
    assert names[0] == "cb_begin" and "cb_admit" in names and "cb_step" in names and "cb_poll" in names and "cb_release" in names
    assert names[-1] == "cb_end"
    admit = next(a for m, a, _ in sent if m == "cb_admit")
    assert admit[0] == [[1, 2]] and admit[1] == [3]


def test_kv_offload_preempts_the_youngest_and_resumes_it():
    """host-DRAM KV tier policy: an admission that finds the pool exhausted preempts the most recently admitted running
    request (swap-out), the new request runs, and the preempted one is swapped back in and finishes with its full output."""
    scripts = {(1,) * 6: list(range(1000, 1300)), (2,) * 6: list(range(2000, 2300)), (3,) * 6: list(range(70, 100))}
    eng = ScriptedEngine(4, scripts, pages=12, step_delay=0.002)      # room for two 6-token prompts, not three
    cb = ContinuousBatcher(eng, steps_per_poll=1, kv_offload=True)

    async def main():
        pad = lambda p: torch.tensor([p])
        a = asyncio.create_task(cb.submit([[1] * 6], pad([1] * 6), 150))
        await asyncio.sleep(0.02)
        b = asyncio.create_task(cb.submit([[2] * 6], pad([2] * 6), 150))
        await asyncio.sleep(0.02)
        c = asyncio.create_task(cb.submit([[3] * 6], pad([3] * 6), 8))          # arrives while both are running
        return await asyncio.gather(a, b, c)
    ra, rb, rc = _run(cb, main())
    assert ra.output_ids[0, 6:].tolist() == list(range(1000, 1150))
    assert rb.output_ids[0, 6:].tolist() == list(range(2000, 2150))      # preempted in the middle, complete in the end
    assert rc.output_ids[0, 6:].tolist() == list(range(70, 78))
    kinds = [c[0] for c in eng.calls]
    assert "swap_out" in kinds and "swap_in" in kinds and kinds.index("swap_out") < kinds.index("swap_in")
    assert cb.stats["preempted"] >= 1 and cb.stats["resumed"] == cb.stats["preempted"]


def test_request_cancelled_while_swapped_out_is_released_not_resumed():
    """a client that disconnects while its request is parked in the host-DRAM tier frees its slot at the next scheduler
    pass; the request is never swapped back in"""
    scripts = {(1,) * 6: list(range(1000, 1400)), (2,) * 6: list(range(2000, 2400)), (3,) * 6: list(range(3000, 3400))}
    eng = ScriptedEngine(4, scripts, pages=12, step_delay=0.002)
    cb = ContinuousBatcher(eng, steps_per_poll=1, kv_offload=True)

    async def main():
        pad = lambda p: torch.tensor([p])
        a = asyncio.create_task(cb.submit([[1] * 6], pad([1] * 6), 120))
        await asyncio.sleep(0.02)
        b = asyncio.create_task(cb.submit([[2] * 6], pad([2] * 6), 300))
        await asyncio.sleep(0.02)
        c = asyncio.create_task(cb.submit([[3] * 6], pad([3] * 6), 120))      # preempts b (the youngest running request)
        for _ in range(200):
            await asyncio.sleep(0.005)
            if any(k[0] == "swap_out" for k in eng.calls):
                break
        b.cancel()
        ra, rc = await asyncio.gather(a, c)
        try:
            await b
        except asyncio.CancelledError:
            pass
        return ra, rc
    ra, rc = _run(cb, main())
    assert ra.output_ids[0, 6:].tolist() == list(range(1000, 1120))
    assert rc.output_ids[0, 6:].tolist() == list(range(3000, 3120))
    kinds = [c[0] for c in eng.calls]
    assert "swap_out" in kinds and "swap_in" not in kinds
    assert cb.stats["cancelled"] == 1 and cb.free_slots == 4 and not eng.slots
