"""Worker for tests/test_tp_gpu.py, launched with torch.distributed.run (one process per GPU): runs the oracle fixtures
through a tensor-parallel engine of WORLD_SIZE ranks and prints, from rank 0, the generated ids plus the teacher-forced
logits error against the fixture (every rank records its vocabulary shard; the shards are gathered over gloo)."""
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
    from tools import synth_weights as W
    cases = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tiny_g4_ids", "tiny_g2_ids", "tiny_moe8_ids"]
    out = {}
    for name in cases:
        c = load_case(name)
        m = c["meta"]
        cfg = W.CONFIGS[m["cfg"]]
        nccl_id = broadcast_nccl_id(rank)   # an ncclUniqueId is single-use: one per communicator
        eng = B200Engine(cfg, max_batch=8, max_seq_len=512, device=int(os.environ["LOCAL_RANK"]),
                         tp_rank=rank, tp_size=world, nccl_id=nccl_id)
        eng.load_weights(W.iter_state_dict(cfg, m["seed"]))
        r = eng.generate(c["input_ids"], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"])
        r2 = eng.generate(c["input_ids"], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"], forced_tokens=c["gen"],
                          want_logits=True)
        shards = [None] * world
        dist.all_gather_object(shards, r2.logits.float())          # [T, B, Vl] per rank, in rank order
        # streaming leader + EOS on the very first token: the leader's callback sees `done` after step 0
        # while the followers have none — all ranks must still run the same number of decode steps
        first = int(c["gen"][0, 0])
        seen = []
        r3 = eng.generate(c["input_ids"][:1], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"], eos_token_ids=[first],
                          forced_tokens=c["gen"][:1], streamer=(lambda step, toks: seen.append(step) and False) if rank == 0 else None)
        r4 = eng.generate(c["input_ids"], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"])   # and the engine still works
        if rank == 0:
            got = torch.cat(shards, 2).permute(1, 0, 2)            # [B, T, V]
            ref = c["step_logits"]
            out[name] = r.output_ids.tolist()
            out[name + ":forced"] = r2.output_ids.tolist()
            out[name + ":max_err"] = float((got - ref).abs().max())
            out[name + ":argmax"] = got.argmax(-1).tolist()
            out[name + ":eos_first_generated"] = r3.num_generated
            out[name + ":again_equal"] = bool(torch.equal(r4.output_ids, r.output_ids))
        eng.close()
        dist.barrier()
    if rank == 0:
        print("TPRESULT " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
