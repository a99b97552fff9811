"""The oracle itself (no GPU): it reproduces its committed fixtures, its streamer restatement matches
transformers' TextIteratorStreamer piece by piece, and BASELINE.json configs[0] — gpt2-small on the CPU backend,
batch 1, 32-token generate through /openai/v1/completions — runs through this repo's server plumbing with the
oracle as the (test-only) backend."""
import asyncio
import os
from threading import Thread

import pytest
import torch
from fastapi.testclient import TestClient

from helpers import GOLDEN, load_case, logits_tol
from oracle import weights as W
from oracle.hf_oracle import OracleGenerativeModel, build_llama


@pytest.mark.parametrize("name", ["tiny_g2_ids", "tiny_g2_stop", "tiny_g2_padinfer", "tiny_g4_peaked_edges"])
def test_oracle_reproduces_fixture(name):
    c = load_case(name)
    m = c["meta"]
    torch.set_num_threads(m.get("threads", 8))
    model = build_llama(W.CONFIGS[m["cfg"]], W.synth_state_dict(W.CONFIGS[m["cfg"]], m["seed"]))
    assert abs(W.checksum(W.synth_state_dict(W.CONFIGS[m["cfg"]], m["seed"])) - m["weights_checksum"]) < 1e-3
    orc = OracleGenerativeModel(model, pad_token_id=m["pad_token_id"])
    r = orc.create_completion(c["input_ids"].tolist(), max_tokens=m["max_tokens"], stop=m["stop"], want_logits=True)
    assert r.finish_reason == m["finish_reason"]
    assert r.prompt_tokens == m["prompt_tokens"] and r.completion_tokens == m["completion_tokens"]
    # bf16 CPU kernels may pick another accumulation order with another thread count / ISA: ids must agree
    # wherever the recorded top-1/top-2 margin is decisive (they agree everywhere on the box that made them)
    tol = logits_tol(c["step_logits"])
    same = r.output_ids == c["output_ids"]
    gen_same = same[:, c["S"]:]
    for b in range(gen_same.shape[0]):
        bad = (~gen_same[b]).nonzero()
        if len(bad):
            assert float(c["margin"][b, int(bad[0])]) <= 2 * tol


def test_text_fixture_pad_resize_and_left_padding():
    """generative_model.py:225-231 + :256-265: left padding, fallback [PAD] resizes the embeddings to len(tokenizer)."""
    from transformers import AutoTokenizer
    c = load_case("tiny_g2_text")
    m = c["meta"]
    tok = AutoTokenizer.from_pretrained(os.path.join(GOLDEN, "byte_tokenizer"))
    model = build_llama(W.CONFIGS[m["cfg"]], W.synth_state_dict(W.CONFIGS[m["cfg"]], m["seed"]))
    orc = OracleGenerativeModel(model, tokenizer=tok)
    assert orc.pad_token_id == m["pad_token_id"] == 256 and model.get_input_embeddings().weight.shape[0] == m["vocab_rows"] == 257
    r = orc.create_completion(m["prompt"], max_tokens=m["max_tokens"])
    assert r.output_ids.shape == c["output_ids"].shape
    assert torch.equal(r.output_ids[:, :c["S"]], c["input_ids"])
    assert (c["input_ids"][3, :-1] == 256).all(), "short prompts are LEFT padded"
    assert r.prompt_tokens == c["S"] * 4, "usage counts pad tokens (q3)"


def test_incremental_detokenizer_matches_text_iterator_streamer():
    from transformers import AutoTokenizer, TextIteratorStreamer
    from kserve_b200.generative_model import IncrementalDetokenizer
    tok = AutoTokenizer.from_pretrained(os.path.join(GOLDEN, "byte_tokenizer"))
    text = "Hello wide world,\nthis is a  streamed\n\nanswer with 中文 and trailing words"
    ids = tok.encode(text)
    ref = TextIteratorStreamer(tok, skip_prompt=False, skip_special_tokens=True)

    def feed():
        for t in ids:
            ref.put(torch.tensor([t]))
        ref.end()
    th = Thread(target=feed)
    th.start()
    want = [p for p in ref if p != ""]
    th.join()
    d = IncrementalDetokenizer(tok)
    got = [p for p in (d.put([t]) for t in ids) if p != ""]
    last = d.end()
    if last != "":
        got.append(last)
    assert got == want and "".join(got) == text


# ---- BASELINE.json configs[0]: gpt2-small, CPU backend, batch 1, 32 tokens, through the server plumbing ----
from kserve_b200.kserve_api import ModelServer  # noqa: E402
from kserve_b200.kserve_api.protocol.rest.openai.openai_chat_adapter_model import OpenAIChatAdapterModel  # noqa: E402
from kserve_b200.kserve_api.protocol.rest.openai.openai_model import ChatPrompt  # noqa: E402
from kserve_b200.kserve_api.protocol.rest.openai.types import Completion, CompletionChoice, UsageInfo  # noqa: E402


class OracleBackedModel(OpenAIChatAdapterModel):
    """TEST-ONLY model: the oracle behind this repo's OpenAI plug-in API (the product class is B200GenerativeModel)."""

    def __init__(self, name, oracle):
        super().__init__(name)
        self.oracle = oracle

    def apply_chat_template(self, request):
        return ChatPrompt(prompt=self.oracle._tokenizer.apply_chat_template(
            [{"role": m.role, "content": m.content} for m in request.messages], tokenize=False, add_generation_prompt=True))

    async def create_completion(self, request, raw_request=None, context=None):
        from kserve_b200.kserve_api.protocol.rest.openai.errors import OpenAIError
        from oracle.hf_oracle import OracleError
        try:
            r = await asyncio.get_running_loop().run_in_executor(None, lambda: self.oracle.create_completion(
                request.prompt, max_tokens=request.max_tokens, stop=request.stop, echo=bool(request.echo)))
        except OracleError as e:       # the reference raises OpenAIError with the same text (generative_model.py:565-572)
            raise OpenAIError(str(e))
        return Completion(id="x", model=request.model,
                          choices=[CompletionChoice(index=i, text=t, finish_reason=r.finish_reason) for i, t in enumerate(r.texts)],
                          usage=UsageInfo(prompt_tokens=r.prompt_tokens, completion_tokens=r.completion_tokens,
                                          total_tokens=r.prompt_tokens + r.completion_tokens))


def test_config1_gpt2_small_cpu_plumbing():
    from transformers import AutoTokenizer, GPT2Config, GPT2LMHeadModel
    torch.manual_seed(0)
    model = GPT2LMHeadModel(GPT2Config(eos_token_id=None, bos_token_id=None)).eval()   # 124 M, fp32, default dims
    tok = AutoTokenizer.from_pretrained(os.path.join(GOLDEN, "byte_tokenizer"))
    orc = OracleGenerativeModel(model, tokenizer=tok, max_length=1024)
    g = torch.Generator().manual_seed(7)
    prompt = torch.randint(3, 250, (1, 16), generator=g).tolist()
    direct = orc.create_completion(prompt, max_tokens=32)
    assert direct.completion_tokens == 32 and direct.finish_reason == "length"
    app = ModelServer().create_application([OracleBackedModel("gpt2", orc)])
    with TestClient(app) as client:
        r = client.post("/openai/v1/completions", json={"model": "gpt2", "prompt": prompt, "max_tokens": 32, "temperature": 0})
        assert r.status_code == 200
        j = r.json()
        assert j["choices"][0]["text"] == direct.texts[0] and j["choices"][0]["finish_reason"] == "length"
        assert j["usage"] == {"prompt_tokens": 16, "completion_tokens": 32, "total_tokens": 48}
        too_long = client.post("/openai/v1/completions", json={"model": "gpt2", "prompt": prompt, "max_tokens": 2000})
        assert too_long.status_code == 500
        assert too_long.json()["error"]["message"] == (
            "This model's maximum context length is 1024 tokens. However, you requested 2016 tokens "
            "(16 in the messages, 2000 in the completion). Please reduce the length of the messages or completion.")
