"""CPU oracle for the Go batcher (pkg/batcher/handler.go) — TEST INFRASTRUCTURE ONLY.

A line-by-line restatement of BatchHandler.batch / batchPredict as a discrete-event simulation: the Go code
is a single goroutine that, per loop iteration, either takes one request from the channel or wakes after
SleepTime (100 us), then evaluates the trigger (handler.go:157-188).  The Go toolchain is absent here, so the
reference cannot be executed; its only golden scenario (handler_test.go:52-174: 10 concurrent clients,
New(32, 50), defaults 32/5000 for New(-1,-1)) is replayed against this restatement in tests/.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

SLEEP_US = 100               # handler.go:34
MAX_BATCH_SIZE = 32          # :35
MAX_LATENCY_MS = 5000        # :36


@dataclass
class Fired:
    at_us: int
    instances: List[Any]
    index: Dict[int, List[int]]          # request id -> indices into instances (InputInfo.Index :50-53)


class GoBatcherOracle:
    def __init__(self, max_batch_size: int, max_latency: int):
        self.MaxBatchSize = MAX_BATCH_SIZE if max_batch_size <= 0 else max_batch_size   # Consume :190-196
        self.MaxLatency = MAX_LATENCY_MS if max_latency <= 0 else max_latency
        self._init(0)

    def _init(self, now_us: int):          # InitializeInfo :89-97
        self.instances: List[Any] = []
        self.context_map: Dict[int, List[int]] = {}
        self.start_us = now_us
        self.current_input_len = 0

    def on_request(self, now_us: int, req_id: int, instances: List[Any]) -> Optional[Fired]:
        """`case req := <-handler.channelIn` (:162-177) followed by the trigger check (:180-186)."""
        if len(self.instances) == 0:
            self.start_us = now_us
        self.current_input_len = len(self.instances)
        self.instances.extend(instances)
        self.context_map[req_id] = [self.current_input_len + i for i in range(len(instances))]
        self.current_input_len = len(self.instances)
        return self.check(now_us)

    def check(self, now_us: int) -> Optional[Fired]:
        elapsed_ms = (now_us - self.start_us) // 1000          # Duration.Milliseconds() truncates
        if self.current_input_len >= self.MaxBatchSize or (elapsed_ms >= self.MaxLatency and self.current_input_len > 0):
            fired = Fired(now_us, self.instances, self.context_map)
            self._init(now_us)
            return fired
        return None

    @staticmethod
    def scatter(fired: Fired, status: int, body: Any, batch_id: str = "uuid") -> Dict[int, dict]:
        """batchPredict :99-155 — what every waiting request receives."""
        out = {}
        if status != 200:
            for rid in fired.index:
                out[rid] = {"message": body if isinstance(body, str) else str(body), "batchId": "", "predictions": None}
            return out
        preds = body.get("predictions") if isinstance(body, dict) else None
        if preds is None or len(preds) != len(fired.instances):
            for rid in fired.index:
                out[rid] = {"message": "size of prediction is not equal to the size of instances", "batchId": batch_id,
                            "predictions": None}
            return out
        for rid, idx in fired.index.items():
            out[rid] = {"message": "", "batchId": batch_id, "predictions": [preds[i] for i in idx]}
        return out


def simulate(arrivals: List[Tuple[int, int, List[Any]]], max_batch_size: int, max_latency: int,
             horizon_us: Optional[int] = None) -> List[Fired]:
    """arrivals: (time_us, request_id, instances) sorted by time.  The goroutine polls every SleepTime when idle."""
    o = GoBatcherOracle(max_batch_size, max_latency)
    fired: List[Fired] = []
    t = 0
    i = 0
    end = horizon_us if horizon_us is not None else (arrivals[-1][0] if arrivals else 0) + (o.MaxLatency + 2) * 1000
    while t <= end:
        if i < len(arrivals) and arrivals[i][0] <= t:
            f = o.on_request(max(t, arrivals[i][0]), arrivals[i][1], arrivals[i][2])
            i += 1
        else:
            nxt = arrivals[i][0] if i < len(arrivals) else end + 1
            t = min(t + SLEEP_US, nxt) if nxt > t else t
            f = o.check(t)
            if i >= len(arrivals) and o.current_input_len == 0:
                if f:
                    fired.append(f)
                break
        if f:
            fired.append(f)
    return fired
