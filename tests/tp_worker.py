"""Worker for tests/test_tp_gpu.py, launched with torch.distributed.run (one process per GPU)."""
import json
import os
import sys

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
    from kserve_b200.engine import B200Engine
    from kserve_b200.tp import broadcast_nccl_id
    from oracle import weights as W
    out = {}
    for name in ("tiny_g4_ids", "tiny_g2_ids", "tiny_moe8_ids"):
        c = load_case(name)
        m = c["meta"]
        nccl_id = broadcast_nccl_id(rank)   # an ncclUniqueId is single-use: one per communicator
        eng = B200Engine(W.CONFIGS[m["cfg"]], max_batch=8, max_seq_len=512, device=int(os.environ["LOCAL_RANK"]),
                         tp_rank=rank, tp_size=world, nccl_id=nccl_id)
        eng.load_weights(W.iter_state_dict(W.CONFIGS[m["cfg"]], m["seed"]))
        r = eng.generate(c["input_ids"], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"])
        r2 = eng.generate(c["input_ids"], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"], forced_tokens=c["gen"])
        out[name] = r.output_ids.tolist()
        out[name + ":forced"] = r2.output_ids.tolist()
        eng.close()
        dist.barrier()
    if rank == 0:
        print("TPRESULT " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
