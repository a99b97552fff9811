"""Worker for tests/test_tp_gpu.py::test_lost_peer_fails_the_call_not_the_context: rank 1 sits out one generate call."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("gloo")
    from helpers import load_case
    from kserve_b200._lib import EngineFault
    from kserve_b200.engine import B200Engine
    from kserve_b200.tp import broadcast_nccl_id
    from tools import synth_weights as W
    c = load_case("tiny_kv8_peaked")
    m = c["meta"]
    cfg = W.CONFIGS[m["cfg"]]
    eng = B200Engine(cfg, max_batch=8, max_seq_len=512, device=int(os.environ["LOCAL_RANK"]), tp_rank=rank, tp_size=world,
                     nccl_id=broadcast_nccl_id(rank))
    eng.load_weights(W.iter_state_dict(cfg, m["seed"]))
    r = eng.generate(c["input_ids"], None, max_new_tokens=4, pad_token_id=m["pad_token_id"])      # both ranks: fine
    ok_first = r.output_ids[:, c["S"]:].tolist() == c["gen"][:, :4].tolist()
    dist.barrier()
    if rank == 0:
        t0 = time.perf_counter()
        try:
            eng.generate(c["input_ids"][:1], None, max_new_tokens=4, pad_token_id=m["pad_token_id"])   # the peer never joins
            outcome = "no error"
        except EngineFault as e:
            outcome = "fault: " + str(e)[:80]
        dt = time.perf_counter() - t0
        x = torch.ones(1024, device="cuda").sum().item()          # the CUDA context is still alive
        try:
            eng.generate(c["input_ids"][:1], None, max_new_tokens=2, pad_token_id=m["pad_token_id"])
            refused = False
        except Exception as e:
            refused = "engine fault" in str(e)
        print(f"TPFAULT first_ok={ok_first} outcome={outcome!r} seconds={dt:.1f} cuda_alive={x == 1024.0} refused_after={refused}")
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
