"""BatchHandler (Python host over the C ABI trigger core) against the scenarios of the reference's
pkg/batcher/handler_test.go, plus the scatter-order assertion the reference lacks."""
import asyncio
import json

import pytest

from kserve_b200.batcher import BatchHandler


def run(coro):
    return asyncio.run(coro)


def test_concurrent_requests_share_one_batch_and_scatter_in_order():
    """handler_test.go:52-90: 10 concurrent clients, New(32, 50): one downstream call, one batchId."""
    calls = []

    async def predictor(path, body):
        calls.append((path, body))
        return 200, {"predictions": [[x[0] * 10] for x in body["instances"]]}   # echo * 10 (proves index mapping)

    async def main():
        h = BatchHandler(32, 50, predictor)
        assert (h.MaxBatchSize, h.MaxLatency) == (32, 50)
        outs = await asyncio.gather(*[h.serve("/v1/models/test:predict", json.dumps({"instances": [[i, i, i]]}).encode())
                                      for i in range(10)])
        return outs
    outs = run(main())
    assert len(calls) == 1 and len(calls[0][1]["instances"]) == 10
    ids = {o[1]["batchId"] for o in outs}
    assert len(ids) == 1 and "" not in ids
    for i, (code, resp) in enumerate(outs):
        assert code == 200 and resp["message"] == "" and resp["predictions"] == [[i * 10]]


def test_batch_size_trigger_counts_instances_and_may_exceed_max():
    """handler.go:181: trigger on instances (not requests); a request is appended whole before the check."""
    sizes = []

    async def predictor(path, body):
        sizes.append(len(body["instances"]))
        return 200, {"predictions": list(range(len(body["instances"])))}

    async def main():
        h = BatchHandler(4, 10_000, predictor)
        a = h.serve("/m:predict", json.dumps({"instances": [1, 2, 3]}).encode())
        b = h.serve("/m:predict", json.dumps({"instances": [4, 5, 6]}).encode())
        return await asyncio.gather(a, b)
    (c1, r1), (c2, r2) = run(main())
    assert sizes == [6]
    assert r1["predictions"] == [0, 1, 2] and r2["predictions"] == [3, 4, 5] and r1["batchId"] == r2["batchId"]


def test_latency_trigger():
    async def predictor(path, body):
        return 200, {"predictions": body["instances"]}

    async def main():
        h = BatchHandler(32, 20, predictor)
        loop = asyncio.get_running_loop()
        t0 = loop.time()
        code, resp = await h.serve("/m:predict", b'{"instances": [7]}')
        return code, resp, loop.time() - t0
    code, resp, dt = run(main())
    assert code == 200 and resp["predictions"] == [7] and 0.015 <= dt < 0.5


def test_downstream_failure_is_broadcast():
    """handler_test.go:93-130 / handler.go:108-117: non-200 -> every waiter gets the body as message, predictions null."""
    async def predictor(path, body):
        return 500, "predictor exploded"

    async def main():
        h = BatchHandler(2, 50, predictor)
        return await asyncio.gather(h.serve("/m:predict", b'{"instances": [1]}'), h.serve("/m:predict", b'{"instances": [2]}'))
    for code, resp in run(main()):
        assert code == 200 and resp == {"message": "predictor exploded", "batchId": "", "predictions": None}


def test_prediction_count_mismatch():
    async def predictor(path, body):
        return 200, {"predictions": [1]}

    async def main():
        h = BatchHandler(2, 50, predictor)
        return await asyncio.gather(h.serve("/m:predict", b'{"instances": [1]}'), h.serve("/m:predict", b'{"instances": [2]}'))
    for code, resp in run(main()):
        assert resp["message"] == "size of prediction is not equal to the size of instances" and resp["predictions"] is None


def test_bad_requests_and_passthrough():
    async def predictor(path, body):
        return 200, {"path": path}

    async def main():
        h = BatchHandler(-1, -1, predictor)
        assert (h.MaxBatchSize, h.MaxLatency) == (32, 5000)          # handler_test.go:133-174
        return (await h.serve("/m:predict", b"{nope"), await h.serve("/m:predict", b'{"instances": []}'),
                await h.serve("/m:predict", b'{"foo": 1}'), await h.serve("/v2/models/m/infer", b'{}'))
    a, b, c, d = run(main())
    assert a == (400, "can't Unmarshal body") and b == (400, "no instances in the request") and c == (400, "no instances in the request")
    assert d == (200, {"path": "/v2/models/m/infer"})               # only :predict is batched (handler.go:224-228)
