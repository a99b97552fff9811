#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its config: output tokens/s for Llama-3-8B (bf16, synthetic weights
of that architecture), 1024-in / 128-out, batch 32, on N GPUs of one node (N>1 = tensor parallel), plus p50 TTFT.

One "step" = one pass of the hot path over one batch: prefill of B x 1024 prompt tokens + 128 greedy tokens.

    python bench.py --gpus 1 --steps 5 --warmup 3              # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 2        # the reference's HF-transformers CPU backend
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # TP=N over NCCL

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLAMA3_8B = dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                 num_attention_heads=32, num_key_value_heads=8, head_dim=128, max_position_embeddings=8192,
                 rms_norm_eps=1e-5, rope_theta=500000.0)


# Mixtral-8x7B dims (8 experts, top-2), run with --model mixtral_8x7b at TP=4
MIXTRAL_8X7B = dict(vocab_size=32000, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                    num_attention_heads=32, num_key_value_heads=8, head_dim=128, max_position_embeddings=8192,
                    rms_norm_eps=1e-5, rope_theta=1000000.0, num_local_experts=8, num_experts_per_tok=2)
MODELS = {"llama3_8b": ("Llama-3-8B", LLAMA3_8B), "mixtral_8x7b": ("Mixtral-8x7B", MIXTRAL_8X7B)}


def workload_config(model_name, B, S, T, world):
    """`config` of BOTH arms (this repo's and --impl reference): the workload the metric is quoted on, plus what one
    step of the reference arm is — stated identically on both sides so the two lines are comparable field by field."""
    return {"workload": f"{model_name} (random-init, bf16) {S}-in/{T}-out batch {B}, greedy",
            "global_batch": B, "prompt_len": S, "gen_len": T, "parallelism": f"tp{world}",
            "reference_sample": f"--impl reference times ONE sequence of this workload per step on the host CPU: batch 1 x "
                                f"{S}-in/{T}-out in full through the restated create_completion (HF transformers CPU backend)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def algorithmic(cfg, B, S_in, S_out, tp=1):
    """SURVEY.md §8(d) figures, computed from the config (per GPU when tp > 1)."""
    H, I, V, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"], cfg["num_hidden_layers"]
    nh, nkv, d = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"]
    E, topk = int(cfg.get("num_local_experts", 0) or 0), int(cfg.get("num_experts_per_tok", 0) or 0)
    attn_params = H * (nh + 2 * nkv) * d + nh * d * H
    # MoE decode streams every expert once B * top_k picks cover them all; flops count the top-k only
    layer_stream = attn_params + (E * 3 * H * I + E * H if E else 3 * H * I)
    layer_params = attn_params + (topk * 3 * H * I + E * H if E else 3 * H * I)
    weight_bytes = (L * layer_stream + V * H) * 2 / tp
    kv_per_tok = L * 2 * nkv * d * 2 / tp
    mean_ctx = S_in + (S_out - 1) / 2 + 1
    decode_bytes = weight_bytes + B * mean_ctx * kv_per_tok
    prefill_flops = (2 * L * layer_params * B * S_in + 2 * S_in * S_in * nh * d * L * B + 2 * V * H * B) / tp
    return dict(weight_bytes=weight_bytes, decode_bytes_per_step=decode_bytes, prefill_flops=prefill_flops,
                kv_bytes_per_token=kv_per_tok, layer_params=layer_params)


def ncu_decode_traffic(cfg, B, S_in, S_out, tp):
    """dram__bytes_read.sum + dram__bytes_write.sum of one decode step, assembled from the committed `ncu --set full`
    captures (profiles/r02_ncu_full_summary.json: one launch per decode kernel at B=32, ctx ~1024; the LM-head launch from
    the round-1 capture of the unchanged kernel) times the launches per step.  Only meaningful for the configuration the
    capture was taken on; None otherwise."""
    if tp != 1 or B != 32 or cfg is not LLAMA3_8B or S_in != 1024:
        return None
    try:
        def gb(e):
            def f(v):
                x, unit = v.split()
                return float(x) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
            return f(e["dram__bytes_read.sum"]) + f(e["dram__bytes_write.sum"])
        rows = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_full_summary.json")))
        r01 = json.load(open(os.path.join(ROOT, "profiles", "r01_final_ncu_full_summary.json")))
        dec = [e for e in rows if e["report"] == "r02_gemm_decode"]
        lm_head = next(gb(e) for e in r01 if e["report"] == "r01f_gemm_decode" and "gemm_tn_kernel<32, 3>" in e["kernel"])
        gate_up = next(gb(e) for e in dec if "gemm_tn_kernel<32, 4>" in e["kernel"])
        parts = sorted(gb(e) for e in dec if "gemm_tn_kernel<32, 5>" in e["kernel"])   # o (34 MB) < qkv (51 MB) < down (118 MB)
        o_proj, qkv, down = parts[0], parts[1], parts[-1]
        attn = next(gb(e) for e in rows if "attn_decode_kernel" in e["kernel"])
        L = cfg["num_hidden_layers"]
        return int(L * (qkv + o_proj + gate_up + down + attn) + lm_head)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation of the path
# ---------------------------------------------------------------------------------------------------
def usable_cores() -> int:
    """CPUs this process may actually run on: affinity mask, capped by the cgroup CPU quota (a container that reports
    128 logical CPUs but is throttled to a few of them makes a 128-thread bf16 matmul ~60x slower, seen on a GPU box)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p_ = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // p_))
        except (OSError, ValueError):
            pass
    return max(1, n)


def pick_cpu_threads() -> int:
    """Thread count for the CPU arm: the fastest of a few candidates on a decode-shaped bf16 GEMV (2 s probe)."""
    import torch
    cap = usable_cores()
    cands = sorted({min(cap, c) for c in (64, 32, 16, 8)}, reverse=True)
    w = torch.randn(14336, 4096).to(torch.bfloat16)
    x = torch.randn(1, 4096).to(torch.bfloat16)
    best, best_t = cands[-1], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        torch.nn.functional.linear(x, w)
        t0 = time.perf_counter()
        for _ in range(8):
            torch.nn.functional.linear(x, w)
        dt = time.perf_counter() - t0
        if dt < best_t * 0.9:       # the larger count stays unless a smaller one is clearly faster
            best, best_t = c, dt
        if dt > 2.0:                # badly oversubscribed: do not try anything larger again
            continue
    torch.set_num_threads(best)
    return best


def cpu_reference(cfg, sample_B, sample_S, sample_T, steps, warmup, budget_s=900.0):
    """HF-transformers CPU backend (what huggingfaceserver runs with --backend huggingface on CPU), through the
    oracle's restated create_completion; random weights of the architecture (values do not affect timing)."""
    import torch
    from oracle.hf_oracle import OracleGenerativeModel
    from transformers import LlamaConfig, LlamaForCausalLM
    pick_cpu_threads()
    hf_cfg = LlamaConfig(**cfg, tie_word_embeddings=False, eos_token_id=None, bos_token_id=None, pad_token_id=None)
    with torch.device("meta"):
        model = LlamaForCausalLM(hf_cfg)
    model = model.to_empty(device="cpu").to(torch.bfloat16)
    g = torch.Generator().manual_seed(0)
    block = (torch.randn(1 << 24, generator=g) * 0.02).to(torch.bfloat16)
    with torch.no_grad():
        for p in model.parameters():
            flat = p.data.view(-1)
            for o in range(0, flat.numel(), block.numel()):
                n = min(block.numel(), flat.numel() - o)
                flat[o:o + n] = block[:n]
            if p.dim() == 1:
                p.data.fill_(1.0)
    for mod in model.modules():
        if hasattr(mod, "inv_freq"):
            inv_freq, scaling = type(mod).compute_default_rope_parameters(mod.config, "cpu")
            mod.register_buffer("inv_freq", inv_freq.float(), persistent=False)
            mod.register_buffer("original_inv_freq", inv_freq.float().clone(), persistent=False)
    model.eval()
    orc = OracleGenerativeModel(model, pad_token_id=cfg["vocab_size"] - 1, max_length=cfg["max_position_embeddings"])
    # The sample is one sequence of the workload IN FULL (sample_S-in / sample_T-out).  Only if the calibration pass
    # (prompt + 1 token, also the first warm-up) projects the whole run past `budget_s` are prompt and generation
    # shortened TOGETHER (same in/out ratio, so the prefill/decode balance of the metric is kept) — and the line says so.
    t0 = time.perf_counter()
    ids = torch.randint(3, 128000, (sample_B, sample_S), generator=g).tolist()
    orc.create_completion(ids, max_tokens=1, temperature=0)
    t_prefill = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.create_completion([row[:64] for row in ids], max_tokens=9, temperature=0)
    t_tok = max(1e-3, (time.perf_counter() - t0) / 9)
    shrink = 1
    while shrink < 8 and (warmup + steps) * (t_prefill / shrink + t_tok * sample_T / shrink) > budget_s:
        shrink *= 2
    full_S, full_T = sample_S, sample_T
    sample_S, sample_T = sample_S // shrink, max(2, sample_T // shrink)
    ids = [row[:sample_S] for row in ids]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = orc.create_completion(ids, max_tokens=sample_T, temperature=0)
        dt = time.perf_counter() - t0
        assert r.completion_tokens == sample_B * sample_T
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    note = "in full" if shrink == 1 else (f"SHORTENED x1/{shrink} from {full_S}-in/{full_T}-out to fit {budget_s:.0f} s of CPU time "
                                          f"(calibration: prefill {t_prefill:.1f} s, {t_tok * 1e3:.0f} ms/token)")
    return dict(value=sample_B * sample_T / (ms / 1e3), ms_per_step=ms, cores=torch.get_num_threads(), shrink=shrink,
                sample=f"Llama-3-8B dims (32 layers, bf16) on the host CPU through the restated create_completion: batch {sample_B} x "
                       f"{sample_S}-in/{sample_T}-out {note}, {steps} timed step(s) after {warmup} warm-up")


def run_reference(args):
    """The reference's own CPU implementation of the path (the restated create_completion around transformers.generate on
    the HF CPU backend — the reference ships no compiled code, oracle/_ref does not apply) on this arm's config / metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model_name, cfg = MODELS[args.model]
    r = cpu_reference(LLAMA3_8B, 1, args.prompt_len, args.gen_len, args.steps, args.warmup, budget_s=args.ref_budget_s)
    line = {
        "impl": "reference", "metric": "output tokens/s", "value": round(r["value"], 3), "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 2),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(model_name, args.batch, args.prompt_len, args.gen_len, args.gpus),
        "cpu_baseline": {"value": round(r["value"], 3), "unit": "tokens/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": round(r["value"], 3), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------
def gpu_weights(cfg, device):
    """random-init weights of the architecture, generated on the device one tensor at a time"""
    import torch
    from kserve_b200.model_spec import llama_tensor_specs as tensor_specs
    g = torch.Generator(device=device).manual_seed(0)
    for name, shape, kind in tensor_specs(cfg):
        if kind == "norm":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        elif kind == "embed":
            t = torch.randn(shape, generator=g, device=device)
        else:
            t = torch.randn(shape, generator=g, device=device) * (shape[-1] ** -0.5)
        yield name, t.to(torch.bfloat16)
        del t


def parity_check(world, rank, local, dev, new_nccl_id):
    """Correctness evidence carried by the bench line itself (every N): the 8-KV-head oracle fixtures (tests/golden,
    generated by oracle/make_goldens.py from the HF CPU backend) through a TP=N engine before anything is timed —
    teacher-forced logits of every step against the fixture (each rank checks its vocabulary shard, max over ranks),
    greedy ids by the margin rule, and the peaked fixture's free-running ids token for token."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from kserve_b200.engine import B200Engine
    from tools import synth_weights as W

    def load(name):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        meta = json.loads(str(z["meta"]))
        out = torch.from_numpy(z["output_ids"].astype(np.int64))
        T = meta["completion_tokens"] // out.shape[0]
        S = out.shape[1] - T
        tv = torch.from_numpy(z["topk_vals"])
        return dict(meta=meta, ids=out[:, :S].contiguous(), gen=out[:, S:].contiguous(), T=T, S=S,
                    logits=torch.from_numpy(z["step_logits"]), margin=tv[..., 0] - tv[..., 1])
    res = {"fixtures": ["tiny_kv8_ids", "tiny_kv8_peaked"], "tp": world}
    ok = True
    for name in res["fixtures"]:
        c = load(name)
        m = c["meta"]
        cfg = W.CONFIGS[m["cfg"]]
        eng = B200Engine(cfg, max_batch=4, max_seq_len=256, device=local, tp_rank=rank, tp_size=world, nccl_id=new_nccl_id())
        eng.load_weights(W.iter_state_dict(cfg, m["seed"]))
        r = eng.generate(c["ids"], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"], forced_tokens=c["gen"], want_logits=True)
        free = eng.generate(c["ids"], None, max_new_tokens=c["T"], pad_token_id=m["pad_token_id"])
        v0, vl = eng.vocab_shard()
        got = r.logits.float().permute(1, 0, 2)                  # [B, T, Vl]: this rank's columns
        ref = c["logits"][..., v0:v0 + vl]
        tol = 4.0 * 2.0 ** -8 * float(c["logits"].abs().max())   # tests/helpers.py TOL_ULPS rule
        t = torch.tensor([float((got - ref).abs().max())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        max_err = float(t.item())
        exact = free.output_ids[:, c["S"]:] == c["gen"]
        decisive = c["margin"] > 2 * tol
        div_ok = all(not bool(decisive[b, int((~exact[b]).nonzero()[0])]) for b in range(exact.shape[0]) if not bool(exact[b].all()))
        if name.endswith("peaked"):
            res["peaked_free_running_exact_match"] = float(exact.float().mean())
            res["peaked_decisive_frac"] = float(decisive.float().mean())
            ok = ok and bool(exact.all())
        else:
            res.update(max_err=round(max_err, 5), tol=round(tol, 5), free_running_token_match=round(float(exact.float().mean()), 4),
                       decisive_steps=int(decisive.sum()), steps=int(decisive.numel()))
            ok = ok and max_err <= tol and div_ok
        eng.close()
        del eng
    res["ok"] = bool(ok)
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    from kserve_b200 import _lib
    from kserve_b200.engine import B200Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def new_nccl_id():
        """one ncclUniqueId per engine (single use), created on rank 0 through the C ABI and broadcast"""
        if world == 1:
            return None
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            import ctypes as C
            raw = C.create_string_buffer(128)
            _lib.check(lib.b200_nccl_unique_id(raw), "nccl_unique_id")
            buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        buf = buf.to(dev)
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().tolist())

    parity = None
    if not args.no_parity_check and 8 % world == 0:
        parity = parity_check(world, rank, local, dev, new_nccl_id)
        torch.cuda.empty_cache()
    nccl_id = new_nccl_id()

    model_name, cfg = MODELS[args.model]
    B, S, T = args.batch, args.prompt_len, args.gen_len
    eng = B200Engine(cfg, max_batch=B, max_seq_len=S + T, max_prefill_tokens=B * S, device=local,
                     tp_rank=rank, tp_size=world, nccl_id=nccl_id)
    eng.load_weights(gpu_weights(cfg, dev))
    torch.cuda.empty_cache()
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(3, min(128000, cfg["vocab_size"] - 8), (B, S), generator=g, dtype=torch.int64).pin_memory()
    pad = cfg["vocab_size"] - 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: prompt staged once, K x (prefill + T-1 decode steps), CUDA events on the
    # engine stream; weights (16 GB) >> L2 (126 MB) so every step re-reads HBM.
    eng.stage(ids, None, max_new_tokens=T, pad_token_id=pad)
    for _ in range(args.warmup):
        eng.run_staged_timed(T - 1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    pre, dec = [], []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        a, b = eng.run_staged_timed(T - 1)
        pre.append(a); dec.append(b)
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = eng.last_launches()
    clocks = sampler.stop() if rank == 0 else None
    out = eng.fetch_staged()
    assert out.shape == (B, S + T), out.shape
    dev_ms = sum(pre) + sum(dec)
    tm = torch.tensor([dev_ms, statistics.median(pre), sum(dec)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms, ttft_p50, dec_ms_total = [float(v) for v in tm.tolist()]
    ms_per_step = dev_ms / args.steps
    value = B * T / (ms_per_step / 1e3)

    # ---- end to end through the public call with HOST buffers (H2D of ids + D2H of results inside)
    e2e_times = []
    for i in range(2 + args.steps):
        barrier()
        t0 = time.perf_counter()
        r = eng.generate(ids, None, max_new_tokens=T, pad_token_id=pad)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= 2:
            e2e_times.append(dt)
    e2e_t = torch.tensor([sum(e2e_times) / len(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = B * T / float(e2e_t.item())
    h2d = 3 * B * S * 4 + 6 * B * 4 + B * eng.max_seq_len // 64 * 4 + 16
    d2h = B * (S + T) * 4 + 16
    # like-for-like companion of the reference arm's sample: ONE sequence of the workload (batch 1 x S-in/T-out) end to
    # end through the same public call — what --impl reference times per step on the host CPU
    one_times = []
    for i in range(2 + min(args.steps, 5)):
        barrier()
        t0 = time.perf_counter()
        eng.generate(ids[:1], None, max_new_tokens=T, pad_token_id=pad)
        torch.cuda.synchronize()
        if i >= 2:
            one_times.append(time.perf_counter() - t0)
    one_t = torch.tensor([sum(one_times) / len(one_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(one_t, op=dist.ReduceOp.MAX)
    one_val = T / float(one_t.item())

    # ---- size-independent properties at the FULL workload size (untimed; every rank takes part in the calls).  The oracle
    # cannot run 32 layers x 32 x 1024 tokens in bench time, so at this size the engine is held to what must be true of
    # ANY correct greedy decoder: the two entry points agree, the batch rows are independent of their position in the
    # batch, and a shorter run is a prefix of a longer one.
    checks = None
    if not args.no_parity_check:
        try:
            full = r.output_ids
            rev = eng.generate(ids.flip(0), None, max_new_tokens=T, pad_token_id=pad).output_ids
            half = eng.generate(ids, None, max_new_tokens=max(1, T // 2), pad_token_id=pad).output_ids
            solo = eng.generate(ids[:1], None, max_new_tokens=T, pad_token_id=pad).output_ids
            checks = {
                "staged_path_equals_generate": bool(torch.equal(out.cpu(), full.cpu())),
                "prompt_echoed": bool(torch.equal(full[:, :S].cpu(), ids)),
                "row_permutation_equivariant": bool(torch.equal(rev.flip(0).cpu(), full.cpu())),
                "shorter_run_is_prefix": bool(torch.equal(half.cpu(), full[:, :half.shape[1]].cpu())),
                # a batch of 1 takes other GEMM tile shapes (block_n 16, other split-K): on random-init weights (near-flat
                # logits) one flipped near-tie changes every later token, so this is reported, not required
                "batch1_row0_token_match": round(float((solo[0, S:].cpu() == full[0, S:].cpu()).float().mean()), 4),
                "tokens_in_vocab": bool(((full >= 0) & (full < cfg["vocab_size"])).all()),
                "size": f"batch {B} x {S}-in/{T}-out, {cfg['num_hidden_layers']} layers",
            }
            checks["ok"] = all(v for k, v in checks.items() if isinstance(v, bool))
        except Exception as e:       # a failed check must not cost the measurement: it is reported instead
            checks = {"ok": False, "error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        return
    peaks = load_peaks()
    alg = algorithmic(cfg, B, S, T, tp=world)
    dec_step_ms = dec_ms_total / (args.steps * (T - 1))
    ach_gbs = alg["decode_bytes_per_step"] / (dec_step_ms * 1e-3) / 1e9
    pre_ms = statistics.median(pre)
    ach_tf = alg["prefill_flops"] / (pre_ms * 1e-3) / 1e12
    line = {
        "metric": "output tokens/s", "value": round(value, 1), "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(model_name, B, S, T, world),
        "notes": {"l2": f"weights {alg['weight_bytes'] / 1e9:.1f} GB per GPU >> 126 MB L2: no flush needed",
                  "timer": "CUDA events on the engine stream, max over ranks"},
        "parity_check": parity,
        "full_size_checks": checks,
        "same_sample_e2e": {"value": round(one_val, 2), "unit": "tokens/s",
                            "sample": f"batch 1 x {S}-in/{T}-out through b200_generate with host buffers: the sample the reference "
                                      "arm times per step (like-for-like numerator for its tokens/s)"},
        "ttft_p50_ms": round(ttft_p50, 2),
        "decode_ms_per_token_step": round(dec_step_ms, 4),
        "wall_s": round(wall, 3),
        "e2e": {"value": round(e2e_val, 1), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "b200_generate (C ABI) with host int64 ids in / host ids out"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "decode step: gemm_tn_kernel<swap-AB> weight streaming + attn_decode KV read",
                     "bound": "hbm", "achieved": round(ach_gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(ach_gbs / peaks["hbm_gbs"], 4), "traffic": ncu_decode_traffic(cfg, B, S, T, world),
                     "traffic_src": "ncu --set full per-launch dram bytes x launches per step (profiles/r02_ncu_full_summary.json); "
                                    "null off the captured configuration", "peak_src": peaks["src"],
                     "algorithmic_bytes_per_launch": int(alg["decode_bytes_per_step"])},
        "roofline_prefill": {"kernel": "prefill: gemm_tn_kernel<256> + attn_prefill", "bound": "tensor",
                             "achieved": round(ach_tf, 1), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                             "frac": round(ach_tf / peaks["tf_sustained"], 4), "peak_src": peaks["src"],
                             "algorithmic_flops_per_launch": alg["prefill_flops"]},
    }
    if world == 1 and not args.no_cpu_baseline and args.model == "llama3_8b":
        eng.close()
        del eng
        torch.cuda.empty_cache()
        r = cpu_reference(cfg, 1, S, T, 2, 1, budget_s=120.0)      # the reference arm's sample: 2 timed steps after 1 warm-up
        line["cpu_baseline"] = {"value": round(r["value"], 3), "unit": "tokens/s", "cores": r["cores"], "kind": "port",
                                "sample": r["sample"]}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The contract is ONE JSON line on stdout: libraries (NCCL prints its version banner to stdout) are kept off it
    by pointing fd 1 at stderr for the whole run and writing the result to the saved descriptor."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="llama3_8b", choices=sorted(MODELS))
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--prompt-len", type=int, default=1024)
    ap.add_argument("--gen-len", type=int, default=128)
    ap.add_argument("--ref-budget-s", type=float, default=900.0,
                    help="--impl reference: CPU seconds the whole --steps/--warmup run may take before the sample is shortened")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
