// Tensor-parallel exchange over NVLink peer memory for the decode step (rows <= 64):
//
//   allreduce_norm_kernel : split-K reduce  +  one-shot all-reduce of the row-parallel GEMM output over peer
//                           loads  +  residual add  +  RMSNorm, in ONE kernel (replaces reduce_partials ->
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
__global__ void __launch_bounds__(kNormThreads)
allreduce_norm_kernel(const P2P pp, bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ xn, float eps,
                      const float* __restrict__ partial, int splits, long long split_stride, long long ld_partial,
                      const bf16* __restrict__ ysrc) {
  TraceScope _ts(TK_RMSNORM);
  pdl_launch_dependents();
  extern __shared__ float row[];  // H floats + 32
  const int H = pp.lay.H;
  float* red = row + H;
  const int r = blockIdx.x;
  const int epoch = pp.row_epoch[r] + 1;   // only this CTA ever writes row_epoch[r]
  const int slot = epoch & 1;
  bf16* ylocal = reinterpret_cast<bf16*>(pp.peer[pp.rank] + pp.lay.y_off) + ((long long)slot * kArRows + r) * H;
  pdl_wait();
  _ts.mark();
  // 1. this rank's contribution, published in its own shared block
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    float f[8];
    if (partial) {
      float a[8];
      sum_partials8(partial + (long long)r * ld_partial + i, splits, split_stride, a);
#pragma unroll
      for (int t = 0; t < 8; ++t) f[t] = a[t];
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
    *reinterpret_cast<uint4*>(ylocal + i) = pack8(f);   // bf16 round: what a non-fused GEMM output would hold
  }
  __threadfence_system();
  __syncthreads();
  // 2. tell every peer that row r of this rank is readable; 3. wait for theirs
  if ((int)threadIdx.x < pp.tp && (int)threadIdx.x != pp.rank) {
    const int peer = threadIdx.x;
    int* remote = reinterpret_cast<int*>(pp.peer[peer] + pp.lay.flag_off) + ((slot * kMaxTp) + pp.rank) * kArRows + r;
    st_sys(remote, epoch);
    const int* mine = reinterpret_cast<const int*>(pp.peer[pp.rank] + pp.lay.flag_off) + ((slot * kMaxTp) + peer) * kArRows + r;
    spin_until(mine, epoch);
    __threadfence_system();
  }
  __syncthreads();
  // 4. sum in rank order (fp32, one rounding), residual add, 5. RMSNorm
  bf16* xr = x + (long long)r * H;
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int rk = 0; rk < pp.tp; ++rk) {
      const bf16* src = reinterpret_cast<const bf16*>(pp.peer[rk] + pp.lay.y_off) + ((long long)slot * kArRows + r) * H + i;
      const uint4 u = ld_sys_v4(src);
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 p2 = unpack_bf16x2(uw[t]);
        acc[2 * t] += p2.x;
        acc[2 * t + 1] += p2.y;
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
  if (threadIdx.x == 0) pp.row_epoch[r] = epoch;
}

}  // namespace b200
