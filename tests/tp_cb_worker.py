"""Worker for tests/test_tp_gpu.py::test_continuous_batching_under_tp: the scheduler runs on rank 0, its b200_cb_* calls
are replicated to the follower ranks (continuous.ReplicatedEngine -> tp.follower_loop)."""
import asyncio
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _Shim:
    """what tp.follower_loop needs from a model"""
    def __init__(self, eng):
        self._engine = eng

    def stop(self):
        pass


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("gloo")
    from helpers import load_case
    from kserve_b200.continuous import ContinuousBatcher
    from kserve_b200.engine import B200Engine
    from kserve_b200.tp import broadcast_nccl_id, follower_loop, leader_call
    from tools import synth_weights as W
    c = load_case("tiny_kv8_peaked")
    m = c["meta"]
    cfg = W.CONFIGS[m["cfg"]]
    eng = B200Engine(cfg, max_batch=8, max_seq_len=512, device=int(os.environ["LOCAL_RANK"]), tp_rank=rank, tp_size=world,
                     nccl_id=broadcast_nccl_id(rank))
    eng.load_weights(W.iter_state_dict(cfg, m["seed"]))
    if rank != 0:
        follower_loop(_Shim(eng))
        eng.close()
        dist.destroy_process_group()
        return
    ids = c["input_ids"]
    prompts = [r.tolist() for r in ids]
    cb = ContinuousBatcher(eng, pad_token_id=0, eos_token_ids=[], steps_per_poll=2, prefill_chunk_tokens=128, prefix_cache=True)
    cb.start()

    async def run():
        t1 = asyncio.create_task(cb.submit([prompts[0]], ids[0:1], c["T"]))
        await asyncio.sleep(0.05)
        t2 = asyncio.create_task(cb.submit(prompts[1:3], ids[1:3], c["T"]))
        t3 = asyncio.create_task(cb.submit([prompts[3]], ids[3:4], 6))
        return await asyncio.gather(t1, t2, t3)
    try:
        r1, r2, r3 = asyncio.run(run())
    finally:
        cb.stop()
    leader_call("stop", (), {})
    S = ids.shape[1]
    print("TPCB " + json.dumps(dict(r1=r1.output_ids[:, S:].tolist(), r2=r2.output_ids[:, S:].tolist(), r3=r3.output_ids[:, S:].tolist())))
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
