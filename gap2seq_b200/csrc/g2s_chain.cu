// g2s_chain.cu - kernels of the chain tier (g2s_chain.cuh): k_chain_walk, one lane per walk; k_chain_finish, one warp
// per gap.
#include <cuda_runtime.h>

#include "g2s_chain.cuh"
#include "g2s_device.h"

namespace g2s {

static const int WALK_THREADS = 64;      // small CTAs: the walks of a pass spread over all SMs
// A warp advances at the pace of its slowest lane's gather, and a gather in a multi-gigabyte table has a long latency
// tail; the walks need no issue bandwidth, so only every walk_stride-th lane carries one (ChainParams.walk_stride,
// default 8: 4 walks per warp).
static const int FINISH_WARPS = 4;

struct ChainWarp {
  static __device__ __forceinline__ int rank() { return threadIdx.x & 31; }
  static __device__ __forceinline__ int size() { return 32; }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
  static __device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
  static __device__ __forceinline__ int min_i32(int v) { return __reduce_min_sync(0xffffffffu, v); }
  static __device__ __forceinline__ int max_i32(int v) { return __reduce_max_sync(0xffffffffu, v); }
  static __device__ __forceinline__ uint32_t add_u32(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }
  static __device__ __forceinline__ unsigned long long bcast_u64(unsigned long long v) { return __shfl_sync(0xffffffffu, v, 0); }
};

// thread 2q walks forward from the left flank of gap order[q], thread 2q + 1 backward from its right flank
__global__ void __launch_bounds__(WALK_THREADS) k_chain_walk(const __grid_constant__ ChainParams P) {
  const uint32_t tt = blockIdx.x * WALK_THREADS + threadIdx.x;
  if (tt % P.walk_stride) return;
  const uint32_t t = tt / P.walk_stride;
  const uint32_t q = t >> 1;
  if (q >= P.n_gaps) return;
  const GapDev gd = P.gaps[P.order[q]];
  const int k = P.table.k;
  uint64_t *W = P.strs + (size_t)q * (P.f_words + P.r_words);
  // ONE call site for both directions: the walks of a warp stay converged (two inlined copies of the loop under an
  // if / else would serialise the forward lanes against the backward lanes)
  const bool fwd = (t & 1) == 0;
  const int maxlev = fwd ? gd.gap_len / 2 + gd.gap_err / 2 + gd.lmf           // the first pruned level (<= Dmax): the levels
                         : gd.rmf + (gd.gap_len + gd.gap_err + 1) / 2;        // beyond it are read off the backward walk (right_half)
  uint64_t x0 = kmer_from_ascii(P.flanks + (fwd ? gd.flank_off : gd.rstart_off + gd.rmf), k);
  // walking predecessors of the last right start = walking successors of its reverse complement
  if (!fwd) x0 = revcomp(x0, k);
  chain_walk(gap_table(P.table, P.tables, P.graph_of, gd.id), x0, maxlev, fwd ? W : W + P.f_words, &P.walks[2 * q + (fwd ? 0 : 1)], !fwd);
}

__global__ void __launch_bounds__(32 * FINISH_WARPS) k_chain_finish(const __grid_constant__ ChainParams P) {
  __shared__ uint64_t tk[FINISH_WARPS][4 * (MAX_FUZ + 1)];
  const uint32_t q = blockIdx.x * FINISH_WARPS + (threadIdx.x >> 5);
  if (q >= P.n_gaps) return;
  const uint32_t gi = P.order[q];
  const GapDev gd = P.gaps[gi];
  const uint64_t *W = P.strs + (size_t)q * (P.f_words + P.r_words);
  chain_finish<ChainWarp>(P, gd, W, W + P.f_words, P.walks[2 * q], P.walks[2 * q + 1], tk[threadIdx.x >> 5], &P.outs[gi]);
}

cudaError_t launch_chain(const ChainParams &p, cudaStream_t st) {
  if (p.n_gaps == 0) return cudaSuccess;
  k_chain_walk<<<(unsigned)(((size_t)2 * p.n_gaps * p.walk_stride + WALK_THREADS - 1) / WALK_THREADS), WALK_THREADS, 0, st>>>(p);
  k_chain_finish<<<(p.n_gaps + FINISH_WARPS - 1) / FINISH_WARPS, 32 * FINISH_WARPS, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace g2s
