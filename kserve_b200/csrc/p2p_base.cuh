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

namespace b200 {

constexpr int kMaxTp = 8;
constexpr int kArRows = 64;   // max decode batch

// Layout of the IPC-shared block (offsets in bytes, computed on the host for a given H)
struct ArLayout {
  size_t y_off, flag_off, cval_off, cidx_off, cflag_off, big_flag_off, total;
  int H;
  __host__ __device__ static ArLayout make(int H) {
    ArLayout l;
    l.H = H;
    l.y_off = 0;                                                   // u64 ll[2][kMaxTp][kArRows][H/2]: (epoch << 32) | bf16x2
    l.flag_off = (size_t)2 * kMaxTp * kArRows * (H / 2) * 8;       // int flag[2][kMaxTp][kArRows] (unused by the LL path)
    l.cval_off = l.flag_off + (size_t)2 * kMaxTp * kArRows * 4;    // float cand_val[2][kMaxTp][kArRows]
    l.cidx_off = l.cval_off + (size_t)2 * kMaxTp * kArRows * 4;    // int cand_idx[2][kMaxTp][kArRows]
    l.cflag_off = l.cidx_off + (size_t)2 * kMaxTp * kArRows * 4;   // int cand_flag[2][kMaxTp][kArRows]
    l.big_flag_off = l.cflag_off + (size_t)2 * kMaxTp * kArRows * 4;   // int big_flag[2 streams][2 phases][kMaxTp] (prefill all-reduce)
    l.total = (l.big_flag_off + (size_t)2 * 2 * kMaxTp * 4 + 255) & ~(size_t)255;   // the prefill exchange buffer starts here
    return l;
  }
};

struct P2P {
  unsigned char* peer[kMaxTp];   // peer[rank] == local block
  int tp, rank;
  ArLayout lay;
  int* row_epoch;                // [0]: decode steps completed (all-reduce epoch base), written by step_update_kernel
  int* cand_epoch;               // [kArRows] local, candidate-exchange epochs
};

__device__ __forceinline__ void st_sys(int* p, int v) { asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_sys_f(float* p, float v) { asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ int ld_sys(const int* p) {
  int v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_sys_v4(const void* p) {
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
// "LL" packets: every 8-byte word carries 4 bytes of payload and the 4-byte epoch, so that the arrival of the
// word itself is the synchronisation (one NVLink one-way store latency, no fence + flag round).  64-bit scalar
// elements are single-copy atomic; the .v2 form only groups two of them into one transaction.
__device__ __forceinline__ void st_ll2(void* p, uint32_t d0, uint32_t d1, uint32_t epoch) {
  const unsigned long long a = ((unsigned long long)epoch << 32) | d0, b = ((unsigned long long)epoch << 32) | d1;
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_ll2(const void* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void spin_until(const int* flag, int want, int code = 1) {
  SpinGuard spins;
  while (ld_sys(flag) != want)
    if (spin_give_up(spins, code)) return;     // recorded in g_fault_code; the host fails the call
}

// push this rank's (max, index) candidate of row b to every rank (including itself); called by one thread
__device__ __forceinline__ void push_candidate(const P2P& pp, int b, float val, int idx) {
  const int epoch = pp.cand_epoch[b] + 1;
  const int slot = epoch & 1;
  const int off = ((slot * kMaxTp) + pp.rank) * kArRows + b;
  for (int peer = 0; peer < pp.tp; ++peer) {
    st_sys_f(reinterpret_cast<float*>(pp.peer[peer] + pp.lay.cval_off) + off, val);
    st_sys(reinterpret_cast<int*>(pp.peer[peer] + pp.lay.cidx_off) + off, idx);
  }
  __threadfence_system();
  for (int peer = 0; peer < pp.tp; ++peer) st_sys(reinterpret_cast<int*>(pp.peer[peer] + pp.lay.cflag_off) + off, epoch);
}

// wait for and merge the candidates of all ranks for row b (lowest index wins ties); called by one thread
__device__ __forceinline__ void merge_candidates(const P2P& pp, int b, float& best, int& tok) {
  const int epoch = pp.cand_epoch[b] + 1;
  const int slot = epoch & 1;
  best = -INFINITY;
  tok = 0x7fffffff;
  for (int rk = 0; rk < pp.tp; ++rk) {
    const int off = ((slot * kMaxTp) + rk) * kArRows + b;
    spin_until(reinterpret_cast<const int*>(pp.peer[pp.rank] + pp.lay.cflag_off) + off, epoch, 3);
    __threadfence_system();
    const float v = ld_sys_f(reinterpret_cast<const float*>(pp.peer[pp.rank] + pp.lay.cval_off) + off);
    const int i = ld_sys(reinterpret_cast<const int*>(pp.peer[pp.rank] + pp.lay.cidx_off) + off);
    if (v > best || (v == best && i < tok)) { best = v; tok = i; }
  }
  pp.cand_epoch[b] = epoch;
}

}  // namespace b200
