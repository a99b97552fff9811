"""Tensor-parallel host plumbing: one process per GPU (torchrun), rank 0 owns the HTTP server; every
generate() call is replicated to the follower ranks over a gloo group so that all ranks enqueue the same
kernels and NCCL collectives in the same order (the engine's own communicator does the per-layer all-reduce)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def broadcast_nccl_id(rank: int) -> bytes:
    """rank 0 creates the ncclUniqueId through the C ABI; everyone receives the 128 bytes."""
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = C.create_string_buffer(128)
        _lib.check(_lib.load().b200_nccl_unique_id(raw), "b200_nccl_unique_id")
        buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        buf = buf.cuda()
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def leader_call(method: str, args: tuple, kwargs: dict) -> None:
    dist.broadcast_object_list([(method, args, kwargs)], src=0)


def follower_loop(model) -> int:
    """Ranks > 0: replay every engine call the leader makes, until it sends 'stop'."""
    while True:
        box = [None]
        dist.broadcast_object_list(box, src=0)
        method, args, kwargs = box[0]
        if method == "stop":
            model.stop()
            return 0
        try:
            getattr(model._engine, method)(*args, **kwargs)
        except Exception as e:   # request-level errors (validation inside the engine) are raised on every rank alike:
            # the leader reports them to the client, the followers must stay in the loop for the next request
            print(f"[kserve_b200 rank {dist.get_rank()}] {method} failed: {e}", flush=True)
