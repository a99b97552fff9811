"""CPU oracle for the KServe LLM predict hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``kserve_b200/`` (the product) may
import, link or execute anything from this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and there only as the checker / timed CPU baseline.

What it restates:

* ``hf_oracle``       – ``HuggingfaceGenerativeModel.create_completion`` /
  ``_handle_request`` (python/huggingfaceserver/huggingfaceserver/
  generative_model.py:286-339, 376-402, 535-646) wrapped around the SAME
  third-party call the reference makes: ``transformers`` ``generate`` on the
  HF CPU backend.  The arithmetic lives in ``transformers``/``torch``, which
  are not vendored under /root/reference (pinned there to transformers 4.57.1
  / torch 2.10.0; installed here: transformers 5.5.0 / torch 2.11.0).
* ``batcher_oracle``  – the Go ``pkg/batcher`` trigger / concat / scatter
  semantics (pkg/batcher/handler.go:99-266).
* ``weights``         – deterministic synthetic checkpoints (no Hub access).

PARITY PIN: the reference's own golden vectors for this path are exact greedy
strings from Hub checkpoints (python/huggingfaceserver/tests/test_model.py:
334-447) which cannot be fetched offline, and the reference packages cannot be
imported here (missing cloudevents/orjson/accelerate/kserve_storage).  The
numerical oracle is therefore pinned by *re-executing the reference's own
dependency call* (transformers.generate) on seeded weights; the protocol-shape
goldens that are usable offline (python/kserve/test/fixtures/openai/*.json,
handler_test.go scenarios) are restated in tests/golden/.  Numerical parity is
"pinned by oracle re-execution, not by reference fixtures".
"""
