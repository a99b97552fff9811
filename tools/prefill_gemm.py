"""Prefill GEMM throughput (TFLOP/s) through b200_op_gemm for the Llama-3-8B layer shapes at T=32768; can load an
alternative library build (B200_LIB_PATH) for A/B on the same box."""
import ctypes as C
import os

import torch

path = os.environ.get("B200_LIB_PATH") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kserve_b200", "lib", "libkserve_b200.so")
lib = C.CDLL(path)
vp, i64 = C.c_void_p, C.c_int64
lib.b200_op_gemm.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64, vp]
lib.b200_last_error.restype = C.c_char_p
T = 32768
dev = "cuda"
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
out = []
for name, N, K, epi in [("qkv", 6144, 4096, 0), ("o+res", 4096, 4096, 1), ("gate_up", 28672, 4096, 2), ("down+res", 4096, 14336, 1)]:
    A = torch.randn(T, K, device=dev, dtype=torch.bfloat16)
    W = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5
    ocols = N // 2 if epi == 2 else N
    O = torch.zeros(T, ocols, device=dev, dtype=torch.bfloat16)
    res = O if epi == 1 else None
    for _ in range(3):
        assert lib.b200_op_gemm(p(A), p(W), p(O), p(res), T, N, K, epi, 256, 1, ocols, None) == 0, lib.b200_last_error()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        lib.b200_op_gemm(p(A), p(W), p(O), p(res), T, N, K, epi, 256, 1, ocols, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out.append(f"{name}: {ms:.3f} ms {2 * T * N * K / ms / 1e9:.0f} TF/s")
print(os.path.basename(path), " | ".join(out))
