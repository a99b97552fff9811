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
        SpinGuard spins;
        while ((uint32_t)(q[rk][0] >> 32) != epoch || (uint32_t)(q[rk][1] >> 32) != epoch ||
               (uint32_t)(q[rk][2] >> 32) != epoch || (uint32_t)(q[rk][3] >> 32) != epoch) {
          if (spin_give_up(spins, 2)) break;
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

// Two-shot variant for tp >= 4 (reduce-scatter + all-gather, both as pushed LL packets, in the same single kernel).
// The one-shot kernel above makes every rank push its whole row to every rank: (tp - 1) * B * H * 4 bytes out per GPU
// (3.6 MB at tp = 8, B = 32 -> 18 us measured, r01).  Here rank j owns columns [j * H/tp, (j+1) * H/tp) of every row:
//   1. scatter : each rank pushes slice j of its own contribution to rank j            ((tp-1)/tp * B * H * 4 B out)
//   2. reduce  : the owner sums the tp contributions of its slice in rank order (fp32, ONE rounding to bf16 — the same
//                value the one-shot kernel computes, so results are bit-identical to it and across ranks)
//   3. gather  : the owner pushes the reduced slice to every rank                      ((tp-1)/tp * B * H * 4 B out)
//   4. every rank polls the complete row, adds the residual and applies the RMSNorm.
// 4x less NVLink traffic than one-shot at tp = 8 for two store latencies instead of one.  Buffers live in the same
// IPC block (y region): scatter [2][tp][kArRows][H/tp/2] u64 at offset 0, gather [2][kArRows][H/2] u64 behind it; the
// epoch-parity double buffering argument of the one-shot kernel carries over (a peer writes epoch e+2 into a slot only
// after it has seen this rank's gather packets of e+1, which this rank sends after it finished reading epoch e).
__global__ void __launch_bounds__(kNormThreads)
allreduce2_norm_kernel(const P2P pp, bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ xn, float eps,
                       const float* __restrict__ partial, int splits, long long split_stride, long long ld_partial,
                       const bf16* __restrict__ ysrc, int ar_index, int ar_per_step) {
  TraceScope _ts(TK_RMSNORM);
  pdl_launch_dependents();
  extern __shared__ float row[];  // H floats + 32
  const int H = pp.lay.H, tp = pp.tp;
  const int Hs = H / tp;          // columns per owner (multiple of 8: H % 64 == 0, tp <= 8)
  float* red = row + H;
  const int r = blockIdx.x;
  const uint32_t epoch = (uint32_t)ld_sys(pp.row_epoch) * (uint32_t)ar_per_step + (uint32_t)ar_index + 1u;
  const int slot = epoch & 1;
  const long long gather_base = (long long)2 * tp * kArRows * (Hs / 2);        // u64 words: scatter region size
  pdl_wait();
  _ts.mark();
  // 1. scatter this rank's contribution: column chunk i goes to its owner
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
    const int owner = i / Hs, c = i - owner * Hs;
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(pp.peer[owner] + pp.lay.y_off) +
                              (((long long)slot * tp + pp.rank) * kArRows + r) * (Hs / 2) + c / 2;
    st_ll2(dst, pk.x, pk.y, epoch);
    st_ll2(dst + 2, pk.z, pk.w, epoch);
  }
  // 2. owner: poll the tp contributions of this rank's slice of row r; one (source, 8-column chunk) per thread task
  const unsigned long long* lbase = reinterpret_cast<const unsigned long long*>(pp.peer[pp.rank] + pp.lay.y_off);
  const int chunks = Hs / 8;
  for (int task = threadIdx.x; task < tp * chunks; task += blockDim.x) {
    const int src_rank = task / chunks, c = (task - src_rank * chunks) * 8;
    const unsigned long long* src = lbase + (((long long)slot * tp + src_rank) * kArRows + r) * (Hs / 2) + c / 2;
    unsigned long long q[4];
    ld_ll2(src, q[0], q[1]);
    ld_ll2(src + 2, q[2], q[3]);
    SpinGuard spins;
    while ((uint32_t)(q[0] >> 32) != epoch || (uint32_t)(q[1] >> 32) != epoch || (uint32_t)(q[2] >> 32) != epoch ||
           (uint32_t)(q[3] >> 32) != epoch) {
      if (spin_give_up(spins, 2)) break;
      ld_ll2(src, q[0], q[1]);
      ld_ll2(src + 2, q[2], q[3]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 p2 = unpack_bf16x2((uint32_t)q[t]);
      row[src_rank * Hs + c + 2 * t] = p2.x;
      row[src_rank * Hs + c + 2 * t + 1] = p2.y;
    }
  }
  __syncthreads();
  // 3. rank-order fp32 sum, one rounding, pushed to every rank's gather buffer (4 columns = one 16-byte LL pair per task)
  for (int c = threadIdx.x * 4; c < Hs; c += blockDim.x * 4) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int sr = 0; sr < tp; ++sr) {
#pragma unroll
      for (int t = 0; t < 4; ++t) a[t] += row[sr * Hs + c + t];
    }
    const uint32_t d0 = pack_bf16x2(a[0], a[1]), d1 = pack_bf16x2(a[2], a[3]);
    const long long goff = gather_base + ((long long)slot * kArRows + r) * (H / 2) + (pp.rank * Hs + c) / 2;
#pragma unroll
    for (int d = 0; d < kMaxTp; ++d) {
      const int peer = (pp.rank + d) % kMaxTp;
      if (peer < tp) st_ll2(reinterpret_cast<unsigned long long*>(pp.peer[peer] + pp.lay.y_off) + goff, d0, d1, epoch);
    }
  }
  __syncthreads();   // `row` is reused below
  // 4. poll the complete reduced row, residual add, RMSNorm
  bf16* xr = x + (long long)r * H;
  const unsigned long long* grow = lbase + gather_base + ((long long)slot * kArRows + r) * (H / 2);
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    unsigned long long q[4];
    ld_ll2(grow + i / 2, q[0], q[1]);
    ld_ll2(grow + i / 2 + 2, q[2], q[3]);
    SpinGuard spins;
    while ((uint32_t)(q[0] >> 32) != epoch || (uint32_t)(q[1] >> 32) != epoch || (uint32_t)(q[2] >> 32) != epoch ||
           (uint32_t)(q[3] >> 32) != epoch) {
      if (spin_give_up(spins, 2)) break;
      ld_ll2(grow + i / 2, q[0], q[1]);
      ld_ll2(grow + i / 2 + 2, q[2], q[3]);
    }
    const uint4 xu = *reinterpret_cast<const uint4*>(xr + i);
    const uint32_t xw[4] = {xu.x, xu.y, xu.z, xu.w};
    float f[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 p2 = unpack_bf16x2(xw[t]);
      float2 y2 = unpack_bf16x2((uint32_t)q[t]);
      f[2 * t] = bf16_round(p2.x + y2.x);
      f[2 * t + 1] = bf16_round(p2.y + y2.y);
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

// ---- prefill: bandwidth-regime all-reduce over peer memory ---------------------------------------------------------
// In place over the row-parallel GEMM output y[T][H] of one micro-batch (up to 256 MiB), which lives inside the
// IPC-shared block of every rank (so peers can address it).  Two-shot by element ranges: rank j owns elements
// [j * n/tp, (j+1) * n/tp): it reads that range from every rank (7 remote NVLink loads + 1 local per 16 bytes), sums in
// rank order in fp32, rounds once and stores the result into EVERY rank's buffer (in place: a range is only ever read by
// its owner).  Same bytes on the wire as NCCL's all-reduce, but:
//   * the CTAs are light (256 threads, <= 128 registers, no shared memory) and co-reside with the persistent 200 KB-smem prefill GEMM CTAs
//     of the OTHER micro-batch, so the exchange really overlaps that GEMM — NCCL's kernels do not fit beside them and
//     only ran in the gaps (r01: TTFT 99.5 ms at TP = 8 for 46 ms of compute);
//   * the sum order is fixed (rank order), every rank holds bit-identical rows.
// Synchronisation: epoch flags in each rank's block.  arrive[rank] = "my y is complete" (the GEMM is an earlier kernel on
// this stream), done[rank] = "my range has been written everywhere"; the kernel's last CTA waits for every done flag, so
// the following rmsnorm kernel sees the complete rows.  `lane` selects one of two flag sets (the two micro-batch streams).
__global__ void __launch_bounds__(256, 2)
allreduce_big_kernel(const P2P pp, long long buf_off, long long elem0, long long n_elems, int epoch, int lane, int* cta_counter) {
  TraceScope _ts(TK_OTHER);
  const int tp = pp.tp;
  int* my_flags = reinterpret_cast<int*>(pp.peer[pp.rank] + pp.lay.big_flag_off) + lane * 2 * kMaxTp;
  if (blockIdx.x == 0 && threadIdx.x < tp) {
    __threadfence_system();
    int* f = reinterpret_cast<int*>(pp.peer[threadIdx.x] + pp.lay.big_flag_off) + lane * 2 * kMaxTp + pp.rank;
    st_sys(f, epoch);
  }
  if (threadIdx.x < tp) spin_until(my_flags + threadIdx.x, epoch);
  __syncthreads();
  const long long vecs = n_elems / 8;                                  // 16-byte vectors (H % 8 == 0)
  const long long per = (vecs + tp - 1) / tp;
  const long long v0 = per * pp.rank, v1 = min(vecs, v0 + per);
  const long long byte0 = buf_off + elem0 * 2;
  // U vectors per thread per pass: tp * U 16-byte loads in flight per thread (U = 4 at tp = 2 ... 1 at tp = 8), enough to
  // cover the ~2 us NVLink round trip at 64 CTAs
  const int U = tp <= 2 ? 4 : (tp <= 4 ? 2 : 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long vb = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; vb < v1; vb += stride * U) {
    uint4 q[4][kMaxTp];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long v = vb + u * stride;
      if (u < U && v < v1) {
#pragma unroll
        for (int r = 0; r < kMaxTp; ++r)
          if (r < tp && (u == 0 || r < 4)) q[u][r] = ld_nc_v4(pp.peer[r] + byte0 + v * 16);   // read once, written before the arrive flags
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long v = vb + u * stride;
      if (u < U && v < v1) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int r = 0; r < kMaxTp; ++r) {
          if (r < tp && (u == 0 || r < 4)) {
            const uint32_t w[4] = {q[u][r].x, q[u][r].y, q[u][r].z, q[u][r].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float2 p2 = unpack_bf16x2(w[t]);
              acc[2 * t] += p2.x;
              acc[2 * t + 1] += p2.y;
            }
          }
        }
        const uint4 o = pack8(acc);
#pragma unroll
        for (int r = 0; r < kMaxTp; ++r)
          if (r < tp)
            *reinterpret_cast<uint4*>(pp.peer[r] + byte0 + v * 16) = o;     // weak stores; __threadfence_system() below orders them before "done"
      }
    }
  }
  // every CTA's stores must be out before this rank says "done"; the last CTA says it and waits for everybody's
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) last = (atomicAdd(cta_counter, 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *cta_counter = 0;
  if (threadIdx.x < tp) {
    __threadfence_system();
    int* f = reinterpret_cast<int*>(pp.peer[threadIdx.x] + pp.lay.big_flag_off) + lane * 2 * kMaxTp + kMaxTp + pp.rank;
    st_sys(f, epoch);
    spin_until(my_flags + kMaxTp + threadIdx.x, epoch);
  }
}

}  // namespace b200
