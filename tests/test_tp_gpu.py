"""Tensor parallel on real GPUs (needs >= 2): TP=2 engine (NCCL all-reduce per row-parallel GEMM, vocab-parallel
LM head) against the oracle fixture with the same margin-aware rule as the single-GPU tests."""
import os

import pytest
import torch

from helpers import load_case, logits_tol

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlap_min_t", ["2048", "1"])   # "1": force the two-micro-batch prefill on the tiny prompts
def test_tp2_matches_oracle_fixture(overlap_min_t):
    import json
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29633", os.path.join(here, "tp_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, B200_PREFILL_OVERLAP_MIN_T=overlap_min_t))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("TPRESULT ")][0]
    ret = {k: torch.tensor(v) for k, v in json.loads(line[len("TPRESULT "):]).items()}
    for name in ("tiny_g4_ids", "tiny_g2_ids", "tiny_moe8_ids"):
        c = load_case(name)
        tol = logits_tol(c["step_logits"])
        assert torch.equal(ret[name + ":forced"], c["output_ids"])
        gen = ret[name][:, c["S"]:]
        for b in range(gen.shape[0]):
            neq = (gen[b] != c["gen"][b]).nonzero()
            if len(neq):
                t = int(neq[0])
                assert float(c["margin"][b, t]) <= 2 * tol, f"{name} row {b} diverges at decisive step {t}"
