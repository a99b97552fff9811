// Batcher: (1) the trigger / index bookkeeping of the Go sidecar's BatchHandler as a deterministic state
// machine (pkg/batcher/handler.go:157-199), (2) device-side request concat (ragged rows -> packed token
// buffers) and response scatter, replacing the JSON concat / loopback HTTP / JSON scatter of
// BatchHandler.batchPredict (handler.go:99-155).
#pragma once
#include <deque>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------
// Trigger state machine.  Time is supplied by the caller (microseconds) so the logic is testable and the
// host language keeps its own concurrency (goroutines + channels in Go, asyncio in Python):
//   add()  == the `case req := <-handler.channelIn` arm      (handler.go:162-177)
//   tick() == the check after every select iteration          (handler.go:180-186)
// A batch fires when CurrentInputLen >= MaxBatchSize (instances, not requests; the whole request is
// appended before the check, so batches may exceed the maximum) or when whole milliseconds since the first
// instance >= MaxLatency and the batch is not empty.
// ---------------------------------------------------------------------------------------------------
struct BatcherCore {
  int max_batch_size, max_latency_ms;
  long long start_us = 0;
  int current_len = 0;
  long long next_ticket = 1;
  struct Req { long long ticket; int first, count; };
  std::vector<Req> reqs;
  std::mutex mu;

  BatcherCore(int mbs, int ml) : max_batch_size(mbs <= 0 ? 32 : mbs), max_latency_ms(ml <= 0 ? 5000 : ml) {}

  long long add(long long now_us, int n_instances) {
    std::lock_guard<std::mutex> g(mu);
    if (current_len == 0) start_us = now_us;  // "if len(Instances) == 0 { Start = now }"
    Req r{next_ticket++, current_len, n_instances};
    reqs.push_back(r);
    current_len += n_instances;
    return r.ticket;
  }
  // returns the number of requests in the fired batch (0 = no trigger)
  int tick(long long now_us, long long* tickets, int* first, int* count, int cap, int* total_instances) {
    std::lock_guard<std::mutex> g(mu);
    const long long elapsed_ms = (now_us - start_us) / 1000;  // Duration.Milliseconds() truncates
    const bool fire = current_len >= max_batch_size || (elapsed_ms >= max_latency_ms && current_len > 0);
    if (!fire) return 0;
    const int n = (int)reqs.size();
    if (n > cap) return -1;
    for (int i = 0; i < n; ++i) { tickets[i] = reqs[i].ticket; first[i] = reqs[i].first; count[i] = reqs[i].count; }
    *total_instances = current_len;
    reqs.clear();
    current_len = 0;
    start_us = now_us;  // InitializeInfo(): Start = now
    return n;
  }
};

// ---------------------------------------------------------------------------------------------------
// Device-side concat.  Input: the raw request rows exactly as they arrived —
//   padded form : ids int64 [B][S] + optional mask int64 [B][S] (left padding)      (OpenAI / V2 leg)
//   ragged form : flat int32 tokens + per-row offsets                               (batcher leg)
// Output: the engine's packed layout (tok, tok_seq, tok_pos, cu_seqlens, last_rows, lens...).  One CTA.
// ---------------------------------------------------------------------------------------------------
struct PackOut {
  int32_t *tok, *tok_seq, *tok_pos, *cu, *seq_slot, *last_rows, *cur_len, *dec_pos, *finished;
};

__global__ void __launch_bounds__(1024)
pack_padded_kernel(const long long* __restrict__ ids, const long long* __restrict__ mask, int B, int S, PackOut o) {
  __shared__ int s_len[64], s_cu[65];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int b = warp; b < B; b += nw) {
    int first = S;
    if (mask) {
      for (int i = lane; i < S; i += 32)
        if (mask[(long long)b * S + i] != 0) { first = i; break; }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
    } else {
      first = 0;
    }
    if (lane == 0) s_len[b] = S - first;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int b = 0; b < B; ++b) { s_cu[b] = t; t += s_len[b]; }
    s_cu[B] = t;
  }
  __syncthreads();
  for (int b = threadIdx.x; b <= B; b += blockDim.x) {
    o.cu[b] = s_cu[b];
    if (b < B) {
      o.seq_slot[b] = b; o.last_rows[b] = s_cu[b + 1] - 1; o.cur_len[b] = s_len[b]; o.dec_pos[b] = s_len[b];
      o.finished[b] = 0;
    }
  }
  for (int b = 0; b < B; ++b) {
    const int len = s_len[b], first = S - len, base = s_cu[b];
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      o.tok[base + i] = (int32_t)ids[(long long)b * S + first + i];
      o.tok_seq[base + i] = b;
      o.tok_pos[base + i] = i;
    }
  }
}

__global__ void __launch_bounds__(1024)
pack_ragged_kernel(const int32_t* __restrict__ flat, const int32_t* __restrict__ offs, int B, PackOut o) {
  for (int b = threadIdx.x; b <= B; b += blockDim.x) {
    o.cu[b] = offs[b];
    if (b < B) {
      const int len = offs[b + 1] - offs[b];
      o.seq_slot[b] = b; o.last_rows[b] = offs[b + 1] - 1; o.cur_len[b] = len; o.dec_pos[b] = len; o.finished[b] = 0;
    }
  }
  for (int b = 0; b < B; ++b) {
    const int base = offs[b], len = offs[b + 1] - base;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      o.tok[base + i] = flat[base + i];
      o.tok_seq[base + i] = b;
      o.tok_pos[base + i] = i;
    }
  }
}

// Response scatter: compacts out_tokens [B][ld] into predictions [B][T] int64 (one row per instance, in
// instance order) so that every waiting request's slice is a contiguous range — the index arithmetic of
// handler.go:139-150 done once on the device instead of per JSON element on the host.
__global__ void scatter_predictions_kernel(const int32_t* __restrict__ out_tokens, int ld, int B, int T,
                                           long long* __restrict__ pred) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < T; i += blockDim.x) pred[(long long)b * T + i] = out_tokens[(long long)b * ld + i];
}

}  // namespace b200
