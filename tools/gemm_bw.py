"""Micro-benchmark: HBM bandwidth of the swap-AB weight-streaming GEMM alone (CUDA events on the launch stream,
rotating over weight copies > L2 so nothing is cache resident)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kserve_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"


def bench(nout, K, batch, epi, bn, splits, reps=20, copies=None):
    nbytes = nout * K * 2
    copies = copies or max(2, int(400e6 // nbytes) + 1)
    Ws = [torch.randn(nout, K, device=dev, dtype=torch.bfloat16) for _ in range(copies)]
    x = torch.randn(max(batch, 64), K, device=dev, dtype=torch.bfloat16)
    if epi == 5:
        out = torch.empty((16, batch, nout), device=dev, dtype=torch.float32)
    elif epi == 4:
        out = torch.empty((batch, nout // 2), device=dev, dtype=torch.bfloat16)
    else:
        out = torch.empty((batch, nout), device=dev, dtype=torch.bfloat16)
    ldo = nout // 2 if epi == 4 else nout
    p = lambda t: C.c_void_p(t.data_ptr())
    for W in Ws:
        rc = lib.b200_op_gemm(p(W), p(x), p(out), None, nout, batch, K, epi, bn, splits, ldo, None)
        assert rc == 0, lib.b200_last_error()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        lib.b200_op_gemm(p(Ws[i % copies]), p(x), p(out), None, nout, batch, K, epi, bn, splits, ldo, None)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    return dict(nout=nout, K=K, batch=batch, epi=epi, bn=bn, splits=splits, us=round(us, 2), GBs=round(nbytes / us / 1e3, 1))


def copy_bw():
    a = torch.empty(1 << 29, device=dev, dtype=torch.bfloat16)
    b = torch.empty_like(a)
    b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    return 2 * a.numel() * 2 * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9


if __name__ == "__main__":
    res = [dict(copy_GBs=round(copy_bw(), 1))]
    B = 32
    for (nout, K, epi, splits) in [(28672, 4096, 4, 1), (28672, 4096, 3, 1), (229376, 512, 3, 1), (57344, 2048, 3, 1),
                                   (4096, 14336, 5, 4), (4096, 14336, 5, 9), (6144, 4096, 5, 3), (4096, 4096, 5, 4),
                                   (128256, 4096, 3, 1), (32768, 4096, 3, 1), (37888, 4096, 3, 1)]:
        res.append(bench(nout, K, B, epi, 32, splits))
    for r in res:
        print(json.dumps(r))
