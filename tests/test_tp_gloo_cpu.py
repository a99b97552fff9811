"""N>1 host-side logic on CPU (gloo, world_size 2): the Megatron split the engine uses for TP — column-parallel
q/k/v and gate/up (with the 16-row gate/up interleave), row-parallel o/down with ONE all-reduce each per layer,
vocab-parallel LM head with a (max, index) merge that breaks ties towards the lowest index — reproduces the
unsharded fp32 forward of the oracle model exactly enough to pick identical tokens."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import weights as W


def shard_layer(sd, l, rank, tp, cfg):
    p = f"model.layers.{l}."
    nh, nkv, d, I = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"], cfg["intermediate_size"]
    ql, kl, Il = nh // tp * d, nkv // tp * d, I // tp
    q = sd[p + "self_attn.q_proj.weight"][rank * ql:(rank + 1) * ql]
    k = sd[p + "self_attn.k_proj.weight"][rank * kl:(rank + 1) * kl]
    v = sd[p + "self_attn.v_proj.weight"][rank * kl:(rank + 1) * kl]
    o = sd[p + "self_attn.o_proj.weight"][:, rank * ql:(rank + 1) * ql]
    g = sd[p + "mlp.gate_proj.weight"][rank * Il:(rank + 1) * Il]
    u = sd[p + "mlp.up_proj.weight"][rank * Il:(rank + 1) * Il]
    dn = sd[p + "mlp.down_proj.weight"][:, rank * Il:(rank + 1) * Il]
    # 16-row gate/up interleave exactly as b200_engine_set_weight lays it out
    gu = torch.empty((2 * Il, g.shape[1]), dtype=g.dtype)
    gu.view(Il // 16, 2, 16, -1)[:, 0] = g.view(Il // 16, 16, -1)
    gu.view(Il // 16, 2, 16, -1)[:, 1] = u.view(Il // 16, 16, -1)
    return q, k, v, o, gu, dn


def rmsnorm(x, w, eps):
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def rope(x, pos, theta):
    d = x.shape[-1]
    inv = 1.0 / (theta ** (torch.arange(0, d, 2).float() / d))
    f = pos[:, None].float() * inv[None]
    c, s = torch.cat([f, f], -1).cos()[:, None], torch.cat([f, f], -1).sin()[:, None]
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    return x * c + torch.cat([-x2, x1], -1) * s


def tp_forward(rank, tp, cfg, sd, ids):
    """last-position greedy token of a [S] prompt with this rank's shards + gloo collectives"""
    H, d, eps = cfg["hidden_size"], cfg["head_dim"], cfg["rms_norm_eps"]
    nhl, nkvl = cfg["num_attention_heads"] // tp, cfg["num_key_value_heads"] // tp
    x = sd["model.embed_tokens.weight"][ids]
    S = x.shape[0]
    pos = torch.arange(S)
    for l in range(cfg["num_hidden_layers"]):
        q, k, v, o, gu, dn = shard_layer(sd, l, rank, tp, cfg)
        xn = rmsnorm(x, sd[f"model.layers.{l}.input_layernorm.weight"], eps)
        Q = rope((xn @ q.T).view(S, nhl, d), pos, cfg["rope_theta"])
        K = rope((xn @ k.T).view(S, nkvl, d), pos, cfg["rope_theta"])
        V = (xn @ v.T).view(S, nkvl, d)
        G = nhl // nkvl
        att = torch.einsum("qhd,khd->hqk", Q, K.repeat_interleave(G, 1)) / d ** 0.5
        att = att.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), 1), float("-inf")).softmax(-1)
        a = torch.einsum("hqk,khd->qhd", att, V.repeat_interleave(G, 1)).reshape(S, nhl * d)
        y = a @ o.T
        dist.all_reduce(y)                       # the per-layer row-parallel reduce #1
        x = x + y
        xn = rmsnorm(x, sd[f"model.layers.{l}.post_attention_layernorm.weight"], eps)
        t = (xn @ gu.T).view(S, -1, 2, 16)
        h = (torch.nn.functional.silu(t[:, :, 0]) * t[:, :, 1]).reshape(S, -1)
        y = h @ dn.T
        dist.all_reduce(y)                       # reduce #2
        x = x + y
    xn = rmsnorm(x[-1:], sd["model.norm.weight"], eps)
    V = cfg["vocab_size"]
    vper = (V + tp - 1) // tp
    v0 = vper * rank
    logits = xn @ sd["lm_head.weight"][v0:v0 + vper].T
    val, idx = logits.max(-1)
    cand = torch.stack([val, (idx + v0).float()], -1)
    allc = [torch.zeros_like(cand) for _ in range(tp)]
    dist.all_gather(allc, cand)
    best_v, best_i = -float("inf"), 1 << 30
    for c in allc:                               # (max, idx) merge, lowest index wins ties — step_update_kernel
        vv, ii = float(c[0, 0]), int(c[0, 1])
        if vv > best_v or (vv == best_v and ii < best_i):
            best_v, best_i = vv, ii
    return best_i


def _worker(rank, tp, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=tp)
    torch.set_num_threads(2)
    cfg = W.CONFIGS["tiny_g2"]
    sd = {k: v.float() for k, v in W.synth_state_dict(cfg, 0).items()}
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(3, 1000, (24,), generator=g)
    tok = tp_forward(rank, tp, cfg, sd, ids)
    if rank == 0:
        ret["tp"] = tok
        from oracle.hf_oracle import OracleGenerativeModel, build_llama
        model = build_llama(cfg, {k: v for k, v in sd.items()}, dtype=torch.float32)
        ref = OracleGenerativeModel(model, pad_token_id=1030).create_completion(ids.tolist(), max_tokens=1)
        ret["ref"] = int(ref.output_ids[0, -1])
    dist.destroy_process_group()


def test_tp2_sharding_matches_unsharded_oracle():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29611, ret), nprocs=2, join=True)
    assert ret["tp"] == ret["ref"]


class _FakeEngine:
    """records the calls replayed on a follower; `generate` with max_new_tokens < 1 fails like the C ABI does"""
    def __init__(self):
        self.calls = []

    def generate(self, ids, mask, **kw):
        if kw.get("max_new_tokens", 1) < 1:
            raise RuntimeError("max_new_tokens must be >= 1")
        self.calls.append(("generate", ids.tolist(), kw))


class _FakeModel:
    def __init__(self):
        self._engine, self.stopped = _FakeEngine(), False

    def stop(self):
        self.stopped = True


def _cmd_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kserve_b200.tp import follower_loop, leader_call
    if rank == 0:
        leader_call("generate", (torch.tensor([[1, 2, 3]]), None), dict(max_new_tokens=4, pad_token_id=0))
        leader_call("generate", (torch.tensor([[7]]), None), dict(max_new_tokens=0))       # fails on every rank alike
        leader_call("generate", (torch.tensor([[9, 9]]), None), dict(max_new_tokens=2))
        leader_call("stop", (), {})
    else:
        m = _FakeModel()
        rc = follower_loop(m)
        ret["calls"] = [(c[0], c[1], c[2]["max_new_tokens"]) for c in m._engine.calls]
        ret["rc"], ret["stopped"] = rc, m.stopped
    dist.destroy_process_group()


def test_follower_replays_leader_calls_and_survives_request_errors():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cmd_worker, args=(2, 29612, ret), nprocs=2, join=True)
    assert ret["calls"] == [("generate", [[1, 2, 3]], 4), ("generate", [[9, 9]], 2)]
    assert ret["rc"] == 0 and ret["stopped"] is True


def _cb_worker(rank, world, port, ret):
    """rank 0 runs the continuous-batching scheduler over a ReplicatedEngine, rank 1 replays its cb_* calls"""
    import asyncio
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kserve_b200.continuous import ContinuousBatcher, ReplicatedEngine
    from kserve_b200.tp import follower_loop, leader_call
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_continuous_cpu import ScriptedEngine
    scripts = {(1, 2): [5, 6, 7, 8], (3,): [9, 2, 4], (4, 4, 4): [11, 12]}
    eng = ScriptedEngine(2, scripts, eos=(2,))
    eng.tp_size = world
    if rank == 0:
        cb = ContinuousBatcher(eng, pad_token_id=0, eos_token_ids=(2,), steps_per_poll=1, prefill_chunk_tokens=128, prefix_cache=True)
        assert isinstance(cb.engine, ReplicatedEngine)

        async def main():
            t = lambda p: torch.tensor([p])
            return await asyncio.gather(cb.submit([[1, 2]], t([1, 2]), 4), cb.submit([[3]], t([3]), 3),
                                        cb.submit([[4, 4, 4]], t([4, 4, 4]), 2))       # three requests over two slots
        cb.start()
        try:
            res = asyncio.run(main())
        finally:
            cb.stop()
        leader_call("stop", (), {})
        ret["leader_out"] = [r.output_ids.tolist() for r in res]
        ret["leader_calls"] = [c for c in eng.calls if c[0] in ("admit", "step", "swap_out", "swap_in")]
        ret["leader_cfg"] = eng.configured
    else:
        m = _FakeModel()
        m._engine = eng
        ret["rc"] = follower_loop(m)
        ret["follower_calls"] = [c for c in eng.calls if c[0] in ("admit", "step", "swap_out", "swap_in")]
        ret["follower_cfg"] = eng.configured
        ret["follower_slots_left"] = len(eng.slots)
    dist.destroy_process_group()


def test_continuous_batcher_calls_are_replicated_to_the_follower_rank():
    """tensor parallel + continuous batching: every engine call the rank-0 scheduler makes (config, admits, steps, polls,
    reads, releases) reaches the follower in the same order, so both ranks hold the same slots / pages at every step"""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cb_worker, args=(2, 29613, ret), nprocs=2, join=True)
    assert ret["leader_out"] == [[[1, 2, 5, 6, 7, 8]], [[3, 9, 2]], [[4, 4, 4, 11, 12]]]
    assert ret["rc"] == 0
    assert ret["leader_cfg"] == ret["follower_cfg"] == (128, True)
    assert ret["leader_calls"] == ret["follower_calls"] and any(c[0] == "admit" for c in ret["follower_calls"])
    assert ret["follower_slots_left"] == 0          # every release was replayed too
