"""Tensor parallel on real GPUs (needs >= 2): TP=2 engine (NCCL all-reduce per row-parallel GEMM, vocab-parallel
LM head) against the oracle fixture with the same margin-aware rule as the single-GPU tests."""
import os

import pytest
import torch

from helpers import load_case, logits_tol

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kserve_b200.engine import B200Engine
    from kserve_b200.tp import broadcast_nccl_id
    from oracle import weights as W
    nccl_id = broadcast_nccl_id(rank)
    for name in ("tiny_g4_ids", "tiny_g2_ids"):
        c = load_case(name)
        m = c["meta"]
        eng = B200Engine(W.CONFIGS[m["cfg"]], max_batch=8, max_seq_len=512, device=rank, tp_rank=rank, tp_size=world, nccl_id=nccl_id)
        eng.load_weights(W.iter_state_dict(W.CONFIGS[m["cfg"]], m["seed"]))
        r = eng.generate(c["input_ids"], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"])
        r2 = eng.generate(c["input_ids"], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"], forced_tokens=c["gen"])
        if rank == 0:
            ret[name] = r.output_ids
            ret[name + ":forced"] = r2.output_ids
        eng.close()
        dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tp2_matches_oracle_fixture():
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, 29633, ret), nprocs=2, join=True)
    for name in ("tiny_g4_ids", "tiny_g2_ids"):
        c = load_case(name)
        tol = logits_tol(c["step_logits"])
        assert torch.equal(ret[name + ":forced"], c["output_ids"])
        gen = ret[name][:, c["S"]:]
        for b in range(gen.shape[0]):
            neq = (gen[b] != c["gen"][b]).nonzero()
            if len(neq):
                t = int(neq[0])
                assert float(c["margin"][b, t]) <= 2 * tol, f"{name} row {b} diverges at decisive step {t}"
