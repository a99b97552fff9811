"""Iteration-level ("continuous") batching scheduler over one B200Engine — SURVEY.md §8(f) rank 1.

The reference executes requests strictly one at a time (`_process_requests`,
python/huggingfaceserver/huggingfaceserver/generative_model.py:341-354; quirk q10) and relies on the Go batcher
(pkg/batcher/handler.go:157-188) to merge concurrent clients into one `generate`.  Here every request joins the
running decode batch at the next step boundary and leaves it as soon as it is finished; the per-request semantics
of the reference are preserved:

  * the rows of ONE request advance in lockstep (they are admitted together) and the request ends when all of its
    rows are finished; finished rows are padded with pad_token_id up to the longest row (utils.py:2796-2797, q3)
  * a stop sequence that matches in ANY row ends the whole request at that step
    (stop_sequence_stopping_criteria.py:36-48, q6) and reports stop_triggered
  * greedy tokens per row are those of `engine.generate` on that request alone.

The scheduler owns the engine: all b200_cb_* calls are made from its thread (the C ABI is not re-entrant).
"""
from __future__ import annotations

import asyncio
import threading
import time
from collections import deque
from dataclasses import dataclass, field
from typing import Callable, Deque, Dict, List, Optional, Sequence

import torch

from .engine import B200Engine, GenerateResult, PoolExhausted


class ReplicatedEngine:
    """Tensor parallel: the scheduler runs on rank 0 only; every b200_cb_* call it makes is first replicated to the
    follower ranks (tp.leader_call -> tp.follower_loop), which execute it on their engine in the same order.  All
    host-side decisions of the engine (slots, pages, prefix hits, chunk boundaries) are deterministic, so the ranks stay
    in lockstep and their poll / read results are identical."""
    _REPLICATED = ("cb_begin", "cb_config", "cb_admit", "cb_step", "cb_poll", "cb_read", "cb_release", "cb_swap_out", "cb_swap_in",
                   "cb_end")

    def __init__(self, engine: B200Engine):
        self._e = engine

    def __getattr__(self, name):
        attr = getattr(self._e, name)
        if name not in self._REPLICATED:
            return attr

        def call(*args, **kwargs):
            from .tp import leader_call
            leader_call(name, args, kwargs)
            return attr(*args, **kwargs)
        return call


@dataclass
class _Request:
    prompts: List[List[int]]                 # un-padded token rows
    padded: torch.Tensor                     # int64 [B, S] as it arrived (echoed in output_ids)
    max_new: int
    stops: List[List[int]]
    pad: int
    done: Callable[[Optional[GenerateResult], Optional[BaseException]], None]
    on_tokens: Optional[Callable[[int, List[int]], None]] = None   # (step, one token per row) for streaming
    sampling: Optional[dict] = None          # repetition_penalty / do_sample / temperature / top_p / top_k / seed (None: greedy)
    slots: List[int] = field(default_factory=list)
    streamed: int = 0
    t_submit: float = field(default_factory=time.perf_counter)
    t_first: float = 0.0
    ttft_recorded: bool = False
    swapped: bool = False                    # preempted to the host-DRAM KV tier
    cancelled: bool = False                  # set from the event loop when the awaiting task went away (client disconnect)


class ContinuousBatcher:
    def __init__(self, engine: B200Engine, pad_token_id: int = 0, eos_token_ids: Sequence[int] = (),
                 steps_per_poll: int = 4, max_prefill_tokens: Optional[int] = None, prefill_chunk_tokens: int = 0,
                 prefix_cache: bool = False, kv_offload: bool = True):
        """prefill_chunk_tokens > 0: admitted prompts are prefilled in chunks of that many packed tokens, one chunk pass
        before every decode step, so a long admit never stalls the running sequences (0: whole prompts at admit).
        prefix_cache: 128-token blocks of earlier prompts stay in the KV pool and are shared by later requests.
        kv_offload: when an admission finds the KV page pool exhausted, the most recently admitted running request is
        preempted to pinned host memory (host-DRAM KV tier) instead of making the new request wait; preempted requests
        are resumed, oldest first, as soon as pages are free again."""
        self.engine = ReplicatedEngine(engine) if getattr(engine, "tp_size", 1) > 1 else engine
        self.prefill_chunk_tokens, self.prefix_cache = int(prefill_chunk_tokens), bool(prefix_cache)
        self.pad = int(pad_token_id or 0)
        self.eos = [int(e) for e in eos_token_ids]
        self.steps_per_poll = max(1, int(steps_per_poll))
        self.max_prefill_tokens = max_prefill_tokens
        self._pending: Deque[_Request] = deque()
        self._running: List[_Request] = []
        self._swapped: Deque[_Request] = deque()     # preempted to host DRAM, resumed oldest first
        self.kv_offload = bool(kv_offload)
        self._cv = threading.Condition()
        self._stop = False
        self._thread: Optional[threading.Thread] = None
        self.free_slots = engine.max_batch
        self.ttft_samples: Deque[float] = deque(maxlen=100000)   # submit -> first token, seconds (load tests / metrics)
        self._pool_blocked_resume = False
        self._pool_blocked = False      # the last admit hit "KV page pool exhausted": retry only after a release
        self.fatal: Optional[BaseException] = None      # set when the scheduler thread died: submit raises it from then on
        self.on_fatal: Optional[Callable[[BaseException], None]] = None   # the model flips `ready` to False here
        # counters for the load tests / metrics
        self.stats: Dict[str, float] = dict(admitted=0, finished=0, cancelled=0, decode_steps=0, row_steps=0, prefill_calls=0)

    # ------------------------------------------------------------------ lifecycle
    def start(self) -> None:
        self.engine.cb_begin(self.pad, self.eos)
        if self.prefill_chunk_tokens or self.prefix_cache:
            self.engine.cb_config(self.prefill_chunk_tokens, self.prefix_cache)
        self._thread = threading.Thread(target=self._loop, name="b200-continuous-batcher", daemon=True)
        self._thread.start()

    def stop(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        if self._thread is not None:
            self._thread.join(timeout=30)
            self._thread = None

    # ------------------------------------------------------------------ submission
    def submit_nowait(self, prompts: List[List[int]], padded: torch.Tensor, max_new_tokens: int,
                      stop_sequences: Sequence[Sequence[int]], done, on_tokens=None, sampling: Optional[dict] = None) -> "_Request":
        stops = [list(map(int, s)) for s in stop_sequences if len(s)]
        if len(prompts) > self.engine.max_batch:
            raise ValueError(f"request of {len(prompts)} prompts exceeds max_batch {self.engine.max_batch}")
        if len(stops) > 4 or any(len(s) > 8 for s in stops):
            raise ValueError("continuous batching supports up to 4 stop sequences of up to 8 tokens each")
        if any(len(p) < 1 or len(p) + int(max_new_tokens) > self.engine.max_seq_len for p in prompts) or int(max_new_tokens) < 1:
            raise ValueError("prompt + max_new_tokens exceeds the engine's max_seq_len (or empty prompt / max_new_tokens < 1)")
        req = _Request(prompts=[list(map(int, p)) for p in prompts], padded=padded, max_new=int(max_new_tokens),
                       stops=stops, pad=self.pad, done=done, on_tokens=on_tokens, sampling=sampling or None)
        with self._cv:
            if self.fatal is not None:
                raise RuntimeError(f"continuous batcher is down: {self.fatal}")
            if self._stop:
                raise RuntimeError("continuous batcher is stopped")
            self._pending.append(req)
            self._cv.notify_all()
        return req

    def cancel(self, req: "_Request") -> None:
        with self._cv:
            req.cancelled = True
            self._cv.notify_all()

    async def submit(self, prompts: List[List[int]], padded: torch.Tensor, max_new_tokens: int,
                     stop_sequences: Sequence[Sequence[int]] = (), on_tokens=None, sampling: Optional[dict] = None) -> GenerateResult:
        loop = asyncio.get_running_loop()
        fut: asyncio.Future = loop.create_future()

        def done(result, err):
            def _set():
                if fut.cancelled():
                    return
                if err is not None:
                    fut.set_exception(err)
                else:
                    fut.set_result(result)
            loop.call_soon_threadsafe(_set)
        req = self.submit_nowait(prompts, padded, max_new_tokens, stop_sequences, done, on_tokens, sampling)
        try:
            return await fut
        except asyncio.CancelledError:
            # the reference wraps its endpoints in `with_cancellation` but cannot abort a running `generate`; here the
            # sequence leaves the batch at the next poll and its slots are reused
            self.cancel(req)
            raise

    # ------------------------------------------------------------------ scheduler thread
    def _admit(self) -> None:
        """Move as many pending requests as fit (slots, prefill token budget) into ONE prefill call."""
        batch: List[_Request] = []
        budget = self.max_prefill_tokens
        free = self.free_slots
        # KV pages the pool can hand out right now: only as many requests as fit are offered to the engine (the head of
        # the queue always is — the engine may still evict cached blocks, and the scheduler may preempt, for it)
        pages_left = None
        if hasattr(self.engine, "cb_stats"):
            try:
                pages_left = int(self.engine.cb_stats()["available_pages"])
            except Exception:
                pages_left = None
        with self._cv:
            while self._pending:
                r = self._pending[0]
                if r.cancelled:
                    self._pending.popleft()
                    self.stats["cancelled"] += 1
                    continue
                need = len(r.prompts)
                toks = sum(len(p) for p in r.prompts)
                if need > free or (budget is not None and batch and toks > budget):
                    break
                if pages_left is not None:
                    pg = sum(-(-(len(p) + r.max_new) // 64) for p in r.prompts)
                    if batch and pg > pages_left:
                        break
                    pages_left -= pg
                self._pending.popleft()
                batch.append(r)
                free -= need
                if budget is not None:
                    budget -= toks
        if not batch:
            return
        while True:
            prompts = [p for r in batch for p in r.prompts]
            max_new = [r.max_new for r in batch for _ in r.prompts]
            stops = [r.stops for r in batch for _ in r.prompts]
            sampling = [r.sampling for r in batch for _ in r.prompts]
            try:
                slots = self.engine.cb_admit(prompts, max_new, stops, sampling if any(sampling) else None)
                break
            except PoolExhausted:            # nothing was admitted
                if len(batch) > 1:           # retry with the head of the queue alone, the others wait their turn
                    with self._cv:
                        for r in reversed(batch[1:]):
                            self._pending.appendleft(r)
                    batch = batch[:1]
                    continue
                if self._running:
                    # host-DRAM KV tier: preempt the most recently admitted running request (its KV pages go to pinned
                    # host memory and back to the pool) and retry; with nothing left to preempt, wait for a release
                    if self.kv_offload and self._preempt_one():
                        continue
                    with self._cv:
                        self._pending.appendleft(batch[0])
                    self._pool_blocked = True
                else:                        # the pool is idle and still too small for this request
                    batch[0].done(None, ValueError("request does not fit the KV page pool"))
                return
            except BaseException as e:       # the whole admit call failed: fail these requests, keep serving
                for r in batch:
                    r.done(None, e)
                return
        self._pool_blocked = False
        now = time.perf_counter()
        i = 0
        for r in batch:
            r.slots = slots[i:i + len(r.prompts)]
            r.t_first = 0.0 if self.prefill_chunk_tokens else now
            i += len(r.prompts)
            self._running.append(r)
        self.free_slots -= len(prompts)
        self.stats["admitted"] += len(batch)
        self.stats["prefill_calls"] += 1

    def _preempt_one(self) -> bool:
        """swap the youngest fully-prefilled running request out to host DRAM; False when there is none"""
        n_gen, fin, _ = self.engine.cb_poll()
        for r in reversed(self._running):
            if r.swapped or r.cancelled or any(n_gen[s] < 1 or fin[s] for s in r.slots):
                continue
            for sl in r.slots:
                self.engine.cb_swap_out(sl)
            r.swapped = True
            self._running.remove(r)
            self._swapped.append(r)
            self.stats["preempted"] = self.stats.get("preempted", 0) + 1
            return True
        return False

    def _drop_cancelled_swapped(self) -> None:
        """a client that went away while its request was parked in the host tier: free the slots, never swap it back in"""
        for r in [r for r in self._swapped if r.cancelled]:
            self._swapped.remove(r)
            for sl in r.slots:
                self.engine.cb_release(sl)
            self.free_slots += len(r.slots)
            self.stats["cancelled"] += 1

    def _resume_swapped(self) -> None:
        """bring preempted requests back (oldest first) while the pool has room"""
        while self._swapped:
            r = self._swapped[0]
            done = []
            try:
                for sl in r.slots:
                    self.engine.cb_swap_in(sl)
                    done.append(sl)
            except PoolExhausted:
                for sl in done:                      # multi-row request only partly back: park it again, try later
                    self.engine.cb_swap_out(sl)
                return
            r.swapped = False
            self._swapped.popleft()
            self._running.append(r)
            self.stats["resumed"] = self.stats.get("resumed", 0) + 1

    def _finish(self, r: _Request, n_out: int, stop: bool) -> None:
        B, S = r.padded.shape
        out = torch.full((B, S + n_out), r.pad, dtype=torch.int64)
        out[:, :S] = r.padded
        for b, sl in enumerate(r.slots):
            toks = self.engine.cb_read(sl, 0, n_out)
            out[b, S:S + len(toks)] = torch.tensor(toks, dtype=torch.int64)
            self.engine.cb_release(sl)
        self.free_slots += len(r.slots)
        self._pool_blocked = self._pool_blocked_resume = False
        self.stats["finished"] += 1
        res = GenerateResult(output_ids=out, stop_triggered=stop, num_generated=n_out, logits=None,
                             prefill_ms=((r.t_first or time.perf_counter()) - r.t_submit) * 1e3, decode_ms=(time.perf_counter() - (r.t_first or r.t_submit)) * 1e3,
                             decode_steps=max(0, n_out - 1), kernel_launches=0)
        r.done(res, None)

    def _collect(self) -> None:
        n_gen, fin, stop = self.engine.cb_poll()
        running, self._running = self._running, []
        ended: List = []
        for r in running:
            if r.cancelled:
                for sl in r.slots:
                    self.engine.cb_release(sl)
                self.free_slots += len(r.slots)
                self._pool_blocked = False
                self.stats["cancelled"] += 1
                continue
            g = [n_gen[s] for s in r.slots]
            if r.t_first == 0.0 and min(g) >= 1:
                r.t_first = time.perf_counter()          # chunked prefill: the first token appears some steps after the admit
            if not r.ttft_recorded and r.t_first:
                r.ttft_recorded = True
                self.ttft_samples.append(r.t_first - r.t_submit)
            stopped = [s for s in r.slots if stop[s]]
            if stopped:
                n_out = min(n_gen[s] for s in stopped)          # lockstep rows: everything after the match is dropped
                end = True
            else:
                n_out = max(g)
                end = all(fin[s] for s in r.slots)
            if r.on_tokens is not None:
                live = [n_gen[s] for s in r.slots if not fin[s]]      # finished rows read as pad from their end on
                upto = n_out if end else min(min(live), n_out) if live else n_out
                if upto > r.streamed:
                    rows = [self.engine.cb_read(s, r.streamed, upto - r.streamed) for s in r.slots]
                    for k in range(upto - r.streamed):
                        r.on_tokens(r.streamed + k, [row[k] if k < len(row) else r.pad for row in rows])
                    r.streamed = upto
            if end:
                ended.append((r, n_out, bool(stopped)))     # off the running list BEFORE done() can be called
            else:
                self._running.append(r)
        for i, (r, n_out, stopped) in enumerate(ended):
            try:
                self._finish(r, n_out, stopped)
            except BaseException as e:
                for r2, _, _ in ended[i:]:                  # not finished yet: fail them once, then let the loop die
                    r2.done(None, e)
                raise

    def _loop(self) -> None:
        try:
            while True:
                with self._cv:
                    while not self._stop and not self._pending and not self._running and not self._swapped:
                        self._cv.wait()
                    if self._stop:
                        break
                if self._swapped:
                    self._drop_cancelled_swapped()
                if self._swapped and not self._pool_blocked_resume:
                    before = len(self._swapped)
                    self._resume_swapped()
                    self._pool_blocked_resume = len(self._swapped) == before and bool(self._running)
                if self._pending and self.free_slots > 0 and not self._pool_blocked and not self._swapped:
                    self._admit()
                if self._running:
                    rows = sum(len(r.slots) for r in self._running)
                    self.engine.cb_step(self.steps_per_poll)
                    self.stats["decode_steps"] += self.steps_per_poll
                    self.stats["row_steps"] += self.steps_per_poll * rows
                    self._collect()
        except BaseException as e:   # engine failure: fail everything that is waiting and refuse new work
            with self._cv:
                self.fatal = e
            if self.on_fatal is not None:
                try:
                    self.on_fatal(e)
                except Exception:
                    pass
        finally:
            # both exits (stop() and a failure) resolve every request still known to the scheduler
            with self._cv:
                waiting = list(self._pending) + self._running + list(self._swapped)
                self._pending.clear()
                self._running = []
                self._swapped.clear()
                err = self.fatal or RuntimeError("continuous batcher stopped")
            for r in waiting:
                try:
                    r.done(None, err)
                except Exception:
                    pass
            try:
                self.engine.cb_end()
            except Exception:
                pass
