"""Prometheus histograms observed inside Model.__call__ and LLM token accounting
(mirrors python/kserve/kserve/metrics.py:17-41)."""
from prometheus_client import Histogram
from pydantic import BaseModel

PROM_LABELS = ["model_name"]
PRE_HIST_TIME = Histogram("request_preprocess_seconds", "pre-process request latency", PROM_LABELS)
POST_HIST_TIME = Histogram("request_postprocess_seconds", "post-process request latency", PROM_LABELS)
PREDICT_HIST_TIME = Histogram("request_predict_seconds", "predict request latency", PROM_LABELS)
EXPLAIN_HIST_TIME = Histogram("request_explain_seconds", "explain request latency", PROM_LABELS)
# additions for the LLM path (tokens/s and TTFT are the headline serving metrics)
TTFT_HIST = Histogram("request_time_to_first_token_seconds", "prefill latency", PROM_LABELS,
                      buckets=(.005, .01, .025, .05, .1, .25, .5, 1, 2.5, 5, 10))
DECODE_TOKENS_PER_S = Histogram("decode_tokens_per_second", "whole-batch decode throughput", PROM_LABELS,
                                buckets=(10, 100, 500, 1000, 2000, 4000, 8000, 16000, 32000))

LLM_STATS_KEY = "llm-stats"


class LLMStats(BaseModel):
    num_prompt_tokens: int = 0
    num_generation_tokens: int = 0


def get_labels(model_name):
    return {PROM_LABELS[0]: model_name}
