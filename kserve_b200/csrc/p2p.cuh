// Tensor-parallel exchange over NVLink peer memory for the decode step (rows <= 64):
//
//   allreduce_norm_kernel : split-K reduce  +  one-shot all-reduce of the row-parallel GEMM output via pushed LL
//                           packets  +  residual add  +  RMSNorm, in ONE kernel (replaces reduce_partials ->
//                           ncclAllReduce -> rmsnorm: 3 launches and a ~10 us collective per row-parallel GEMM).
//   candidate exchange    : vocab-parallel greedy argmax candidates are pushed to every peer by argmax_kernel and
//                           merged by step_update_kernel (replaces 2 ncclAllGather per step).
//
// Every rank owns one `ArShared` block (cudaMalloc, exported with cudaIpcGetMemHandle); peers map it with
// cudaIpcOpenMemHandle.  Synchronisation is by epoch-valued flags written with system-scope stores into the
// *reader's* block (so readers spin on local memory), double buffered so that one flag round per all-reduce
// suffices: a rank can be at most one all-reduce ahead of a peer, and it then uses the other slot.
// The sum is taken in rank order in fp32 and rounded once, so every rank computes bit-identical rows.
#pragma once
#include "common.cuh"
#include "ops.cuh"
#include "p2p_base.cuh"

namespace b200 {

// x[r] += allreduce(y_local[r]);  xn[r] = rmsnorm(x[r]) * w      (one CTA of 512 threads per row)
// y_local is either the sum of this rank's split-K partials (partial != null) or a bf16 matrix (ysrc).
//
// Exchange: every rank PUSHES its bf16 row into slot [epoch & 1][rank][r] of every rank's block (its own
// included) as LL packets (payload + epoch in one 8-byte word), then polls its own block until the packets of all
// ranks carry the current epoch.  Cost: one NVLink store latency; no fence, no flag round, no remote loads.
// Double buffering is sufficient: a rank pushes epoch e+2 only after it has seen every peer's e+1 packets, i.e.
// after every peer has finished reading epoch e (its e+1 kernel pushes after pdl_wait(), when its e kernel is done).
__global__ void __launch_bounds__(kNormThreads)
allreduce_norm_kernel(const P2P pp, bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ xn, float eps,
                      const float* __restrict__ partial, int splits, long long split_stride, long long ld_partial,
                      const bf16* __restrict__ ysrc, int ar_index, int ar_per_step) {
  TraceScope _ts(TK_RMSNORM);
  pdl_launch_dependents();
  extern __shared__ float row[];  // H floats + 32
  const int H = pp.lay.H;
  float* red = row + H;
  const int r = blockIdx.x;
  // Epoch = (decode steps so far) * (all-reduces per step) + index of this all-reduce in the step + 1.  The step
  // counter is only written by step_update_kernel at the end of the previous step's graph, so reading it before
  // pdl_wait() is safe however far ahead the programmatic launches run.  ar_per_step is even: slots alternate.
  const uint32_t epoch = (uint32_t)ld_sys(pp.row_epoch) * (uint32_t)ar_per_step + (uint32_t)ar_index + 1u;
  const int slot = epoch & 1;
  // u64 index of (src rank, row r, word pair 0) inside any rank's block
  const long long mine_off = (((long long)slot * kMaxTp + pp.rank) * kArRows + r) * (H / 2);
  pdl_wait();
  _ts.mark();
  // 1. this rank's contribution (bf16 rounded: what a non-fused GEMM output would hold), pushed to every rank
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    float f[8];
    if (partial) {
      sum_partials8(partial + (long long)r * ld_partial + i, splits, split_stride, f);
    } else {
      uint4 u = *reinterpret_cast<const uint4*>(ysrc + (long long)r * H + i);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 p2 = unpack_bf16x2(uw[t]);
        f[2 * t] = p2.x;
        f[2 * t + 1] = p2.y;
      }
    }
    const uint4 pk = pack8(f);
#pragma unroll
    for (int d = 0; d < kMaxTp; ++d) {
      const int peer = (pp.rank + d) % kMaxTp;   // start with the own block, spread the first remote target
      if (peer < pp.tp) {
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(pp.peer[peer] + pp.lay.y_off) + mine_off + i / 2;
        st_ll2(dst, pk.x, pk.y, epoch);
        st_ll2(dst + 2, pk.z, pk.w, epoch);
      }
    }
  }
  // 2. poll the local block for every rank's packets, sum in rank order (fp32, one rounding), residual add
  bf16* xr = x + (long long)r * H;
  const unsigned long long* base = reinterpret_cast<const unsigned long long*>(pp.peer[pp.rank] + pp.lay.y_off);
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    unsigned long long q[kMaxTp][4];
#pragma unroll
    for (int rk = 0; rk < kMaxTp; ++rk) {
      if (rk < pp.tp) {
        const unsigned long long* src = base + (((long long)slot * kMaxTp + rk) * kArRows + r) * (H / 2) + i / 2;
        ld_ll2(src, q[rk][0], q[rk][1]);
        ld_ll2(src + 2, q[rk][2], q[rk][3]);
      }
    }
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int rk = 0; rk < kMaxTp; ++rk) {
      if (rk < pp.tp) {
        const unsigned long long* src = base + (((long long)slot * kMaxTp + rk) * kArRows + r) * (H / 2) + i / 2;
        uint32_t spins = 0;
        while ((uint32_t)(q[rk][0] >> 32) != epoch || (uint32_t)(q[rk][1] >> 32) != epoch ||
               (uint32_t)(q[rk][2] >> 32) != epoch || (uint32_t)(q[rk][3] >> 32) != epoch) {
          if (++spins > (1u << 26)) {
            printf("b200: peer packet timeout (row %d thread %d rank %d epoch %u)\n", r, threadIdx.x, rk, epoch);
            __trap();
          }
          ld_ll2(src, q[rk][0], q[rk][1]);
          ld_ll2(src + 2, q[rk][2], q[rk][3]);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float2 p2 = unpack_bf16x2((uint32_t)q[rk][t]);
          acc[2 * t] += p2.x;
          acc[2 * t + 1] += p2.y;
        }
      }
    }
    const uint4 xu = *reinterpret_cast<const uint4*>(xr + i);
    const uint32_t xw[4] = {xu.x, xu.y, xu.z, xu.w};
    float f[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 p2 = unpack_bf16x2(xw[t]);
      f[2 * t] = bf16_round(p2.x + bf16_round(acc[2 * t]));
      f[2 * t + 1] = bf16_round(p2.y + bf16_round(acc[2 * t + 1]));
    }
    *reinterpret_cast<uint4*>(xr + i) = pack8(f);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      row[i + t] = f[t];
      ss += f[t] * f[t];
    }
  }
  const float tot = block_sum(ss, red);
  const float rs = 1.0f / sqrtf(tot / (float)H + eps);
  bf16* o = xn + (long long)r * H;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    const uint4 wu = *reinterpret_cast<const uint4*>(w + i);
    const uint32_t ww[4] = {wu.x, wu.y, wu.z, wu.w};
    float g[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 w2 = unpack_bf16x2(ww[t]);
      g[2 * t] = w2.x * bf16_round(row[i + 2 * t] * rs);
      g[2 * t + 1] = w2.y * bf16_round(row[i + 2 * t + 1] * rs);
    }
    *reinterpret_cast<uint4*>(o + i) = pack8(g);
  }
}

}  // namespace b200
