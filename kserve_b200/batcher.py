"""BatchHandler — the Go agent's request batcher (pkg/batcher/handler.go) for hosts without the sidecar.

Same constructor meaning (`New(maxBatchSize, maxLatency, handler, logger)` handler.go:210-220; <= 0 selects the
defaults 32 / 5000 ms), same wire types (`{"instances": [...]}` in, `{"message", "batchId", "predictions"}` out,
handler.go:39-59), same error behaviour (400 "can't Unmarshal body" / "no instances in the request"
:233-243; downstream failure -> 200 with `message` set and `predictions: null` :108-117; prediction-count
mismatch :130-137).  The trigger decision and the per-request index bookkeeping live in the C ABI state
machine (b200_batcher_add / b200_batcher_tick); this class only supplies the clock and the awaiting.
"""
from __future__ import annotations

import asyncio
import ctypes as C
import json
import re
import time
import uuid
from typing import Any, Awaitable, Callable, Dict, List, Optional, Tuple

from . import _lib

SLEEP_TIME_S = 100e-6          # handler.go:34
PREDICT_VERB = re.compile(r":predict$")


class BatchHandler:
    def __init__(self, max_batch_size: int, max_latency: int,
                 next_handler: Callable[[str, Dict[str, Any]], Awaitable[Tuple[int, Any]]], logger=None,
                 clock: Callable[[], float] = time.monotonic):
        """next_handler(path, {"instances": [...]}) -> (status_code, body) where body is a dict with
        "predictions" on success (anything else on failure), i.e. the downstream predictor."""
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.b200_batcher_create(max_batch_size, max_latency, C.byref(h)), "b200_batcher_create")
        self.h = h
        a, b = C.c_int32(), C.c_int32()
        self.lib.b200_batcher_config(self.h, C.byref(a), C.byref(b))
        self.MaxBatchSize, self.MaxLatency = a.value, b.value
        self.next, self.log, self.clock = next_handler, logger, clock
        self._instances: List[Any] = []
        self._waiters: Dict[int, asyncio.Future] = {}
        self._path = ""
        self._timer: Optional[asyncio.Task] = None
        self._lock = asyncio.Lock()

    def __del__(self):
        try:
            self.lib.b200_batcher_destroy(self.h)
        except Exception:
            pass

    def _now_us(self) -> int:
        return int(self.clock() * 1e6)

    async def _tick(self):
        cap = max(1, len(self._waiters))
        tickets = (C.c_int64 * cap)()
        first = (C.c_int32 * cap)()
        count = (C.c_int32 * cap)()
        n, total = C.c_int32(0), C.c_int32(0)
        _lib.check(self.lib.b200_batcher_tick(self.h, self._now_us(), cap, tickets, first, count, C.byref(n),
                                              C.byref(total)), "b200_batcher_tick")
        if n.value == 0:
            return
        instances, self._instances = self._instances, []
        waiters = [(self._waiters.pop(tickets[i]), first[i], count[i]) for i in range(n.value)]
        path = self._path
        await self._batch_predict(path, instances, waiters)

    @staticmethod
    def _resolve(fut: asyncio.Future, value: Dict[str, Any]) -> None:
        # a waiter whose client went away is cancelled (its serve() task was): skip it, the others must still be answered
        if not fut.done():
            fut.set_result(value)

    async def _batch_predict(self, path, instances, waiters):
        """handler.go:99-155.  Every waiter of the batch is resolved on every path out of this function (a waiter popped
        from the table and never answered would hang its client forever)."""
        def answer_all(message: str, batch_id: str) -> None:
            for fut, _, _ in waiters:
                self._resolve(fut, {"message": message, "batchId": batch_id, "predictions": None})
        try:
            try:
                code, body = await self.next(path, {"instances": instances})
            except asyncio.CancelledError:
                raise
            except Exception as e:  # a raising predictor is a non-200 downstream
                code, body = 500, str(e)
            if code != 200:
                answer_all(body if isinstance(body, str) else json.dumps(body), "")
                return
            batch_id = str(uuid.uuid4())
            preds = body.get("predictions") if isinstance(body, dict) else None
            if not isinstance(preds, list):
                answer_all("can't Unmarshal predictions", batch_id)
                return
            if len(preds) != len(instances):
                answer_all("size of prediction is not equal to the size of instances", batch_id)
                return
            for fut, f, c in waiters:
                self._resolve(fut, {"message": "", "batchId": batch_id, "predictions": preds[f:f + c]})
        finally:
            for fut, _, _ in waiters:       # cancellation / unexpected error inside the block above
                if not fut.done():
                    fut.set_result({"message": "batch aborted", "batchId": "", "predictions": None})

    async def _latency_timer(self):
        # stands in for the `case <-time.After(SleepTime)` arm: re-check until the waiting batch has fired
        while self._waiters:
            await asyncio.sleep(max(SLEEP_TIME_S, min(0.001, self.MaxLatency / 1e3)))
            async with self._lock:
                try:
                    await self._tick()
                except Exception as e:   # the timer must outlive a failing batch: later waiters depend on it
                    if self.log:
                        self.log.error("batcher tick failed: %s", e)

    async def serve(self, path: str, body: bytes) -> Tuple[int, Any]:
        """handler.go:222-266 -> (status, response).  Non-:predict paths go straight to the next handler."""
        if not PREDICT_VERB.search(path):
            try:
                return await self.next(path, json.loads(body) if body else {})
            except json.JSONDecodeError:
                return 400, "can't Unmarshal body"
        try:
            req = json.loads(body)
        except (json.JSONDecodeError, UnicodeDecodeError):
            return 400, "can't Unmarshal body"
        instances = req.get("instances") if isinstance(req, dict) else None
        if not isinstance(instances, list) or len(instances) == 0:
            return 400, "no instances in the request"
        fut: asyncio.Future = asyncio.get_running_loop().create_future()
        async with self._lock:
            ticket = C.c_int64(0)
            _lib.check(self.lib.b200_batcher_add(self.h, self._now_us(), len(instances), C.byref(ticket)), "b200_batcher_add")
            self._waiters[ticket.value] = fut
            self._instances.extend(instances)
            self._path = path
            await self._tick()
            if self._waiters and (self._timer is None or self._timer.done()):
                self._timer = asyncio.ensure_future(self._latency_timer())
        return 200, await fut
