// Weight-stream L2 prefetcher for the decode step.
//
// A decode step is one long dependent chain of small kernels (7 per layer); each weight-streaming GEMM can only pull
// its weights while it runs, so HBM idles in every gap of the chain (kernel tails, griddepcontrol waits, the epilogue ->
// next prologue latency): r01 measured 0.69 of the HBM peak with DRAM traffic == algorithmic bytes, i.e. pure latency.
// The weights, however, depend on nothing.  This kernel runs on a parallel branch of the step's CUDA graph for the whole
// step, one tiny CTA per SM, and walks the SAME per-CTA work lists as the GEMM kernels (gemm_get_seg) in launch order,
// issuing `cp.async.bulk.prefetch.tensor.2d.L2` for exactly the tiles the GEMM CTA of the same index will load, a
// bounded number of k-blocks (`lead`) ahead of what that CTA has published in pf_prog.  The 126 MB L2 is the staging
// buffer: HBM keeps streaming through the stalls of the chain, and the GEMMs find their tiles in L2.
//
// Pacing / safety: never more than `lead` k-blocks (16 KB each) per CTA ahead (148 x lead x 16 KB of L2); units the
// consumer has already passed are skipped, never re-fetched; a consumer that does not move for `timeout_ns` makes the CTA
// exit (prefetching is an optimisation, it must never be able to hang the step).
#pragma once
#include "gemm_tcgen05.cuh"

namespace b200 {

struct alignas(64) PfGemm {      // one weight-streaming GEMM of the step, in launch order
  CUtensorMap tmap_a;            // the weight operand's tensor map (128 rows x 64 columns boxes)
  GemmSched sched;               // 6 ints
  int grid;                      // CTAs of the GEMM launch
  int pad[9];
};
static_assert(sizeof(PfGemm) % 64 == 0, "tensor maps in the table must stay 64-byte aligned");

constexpr int kPfMaxGemms = 320;

__global__ void __launch_bounds__(32) weight_prefetch_kernel(const PfGemm* __restrict__ tab, int n_gemms,
                                                             const unsigned long long* __restrict__ prog,
                                                             const unsigned int* __restrict__ seq_ptr, int lead,
                                                             unsigned long long timeout_ns, int mode, int skip) {
  // skip: the consumer publishes the k-blocks whose MMAs were issued; its own TMA ring is up to `skip` k-blocks ahead of that,
  // so units closer than `skip` to the published position have been requested already and are not prefetched again
  // mode (A/B diagnostics): 0 normal, 1 dry run (walk + pace, no prefetch instruction), 2 exit at once
  TraceScope _ts(11);
  __shared__ int cum[kPfMaxGemms + 1];   // units of this CTA before GEMM g
  if (threadIdx.x != 0 || mode == 2) return;
  const int c = blockIdx.x;
  const unsigned int seq = *reinterpret_cast<const volatile unsigned int*>(seq_ptr);
  const volatile unsigned long long* slot = prog + c;
  int issued = 0;          // units of this CTA's stream handled so far (prefetched or skipped)
  int consumed = 0;
  unsigned long long last_move = global_timer();
  int last_consumed = -1;
  for (int g = 0; g < n_gemms; ++g) {
    cum[g] = issued;
    const PfGemm& G = tab[g];
    if (c >= G.grid) continue;
    GemmSeg sg;
    for (int idx = 0; gemm_get_seg(G.sched, idx, sg, G.sched.m_tiles, c, G.grid); ++idx) {
      for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
        // pacing: re-read the consumer's position when we are at the edge of the window
        while (issued - consumed >= lead) {
          const unsigned long long v = *slot;
          int pos = 0;
          if ((unsigned int)(v >> 32) == seq) {
            const int pg = (int)((v >> 16) & 0xffff);
            pos = (pg <= g ? cum[pg] : issued) + (int)(v & 0xffff);
          }
          consumed = pos;
          if (issued - consumed < lead) break;
          const unsigned long long now = global_timer();
          if (consumed != last_consumed) { last_consumed = consumed; last_move = now; }
          else if (now - last_move > timeout_ns) return;     // the chain is not running (aborted / different path): give up
          __nanosleep(200);
        }
        if (issued >= consumed + skip && mode == 0)      // the consumer has not requested this unit yet
          tma_prefetch_l2_2d(&G.tmap_a, kb * kGemmBlockK, sg.m_t * kGemmBlockM);
        ++issued;
      }
    }
  }
}

}  // namespace b200
