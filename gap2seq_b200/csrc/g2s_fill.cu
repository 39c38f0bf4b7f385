// g2s_fill.cu - the persistent fill kernels.  Each thread group (a warp in the thin tier, a 512-thread CTA in
// the wide tier) owns one scratch slice, pulls gaps from a device work queue (longest first) and runs
// process_gap (g2s_fill.cuh) on them, so many short gaps share one launch.
#include <cuda_runtime.h>

#include "g2s_device.h"
#include "g2s_fill.cuh"

namespace g2s {

static const int WARPS_PER_CTA = 1;   // one warp group per CTA: few long gaps spread over all SMs
static const int CTA_THREADS = 512;

struct WarpGroup {
  static constexpr bool kWarp = true;
  static __device__ __forceinline__ int rank() { return threadIdx.x & 31; }
  static __device__ __forceinline__ int size() { return 32; }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
};
struct BlockGroup {
  static constexpr bool kWarp = false;
  static __device__ __forceinline__ int rank() { return threadIdx.x; }
  static __device__ __forceinline__ int size() { return blockDim.x; }
  static __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <class G>
__device__ __forceinline__ void group_loop(const FillParams &P, GShared &sh, uint32_t group_id) {
  const Slice S = make_slice(P.scratch + (size_t)group_id * P.lay.bytes, P.lay);
  for (;;) {
    G::sync();
    if (G::rank() == 0) sh.queue_item = atomicAdd(P.queue_head, 1u);
    G::sync();
    const uint32_t q = ld_sh(&sh.queue_item);
    if (q >= P.n_gaps) break;
    const uint32_t gi = P.order[q];
    const GapDev gd = P.gaps[gi];
    process_gap<G>(P, gd, S, sh, &P.outs[gi]);
  }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_fill_warp(const __grid_constant__ FillParams P, int groups) {
  __shared__ GShared sh[WARPS_PER_CTA];
  const uint32_t w = threadIdx.x >> 5;
  const uint32_t group_id = blockIdx.x * WARPS_PER_CTA + w;
  if (group_id >= (uint32_t)groups) return;
  group_loop<WarpGroup>(P, sh[w], group_id);
}

__global__ void __launch_bounds__(CTA_THREADS) k_fill_cta(const __grid_constant__ FillParams P, int groups) {
  __shared__ GShared sh;
  if (blockIdx.x >= (uint32_t)groups) return;
  group_loop<BlockGroup>(P, sh, blockIdx.x);
}

int fill_warps_per_cta() { return WARPS_PER_CTA; }

int fill_max_resident_warp_groups(int sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fill_warp, WARPS_PER_CTA * 32, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return sms * per_sm * WARPS_PER_CTA;
}

int fill_max_resident_cta_groups(int sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fill_cta, CTA_THREADS, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return sms * per_sm;
}

cudaError_t launch_fill_warp(const FillParams &p, int groups, cudaStream_t st) {
  const int ctas = (groups + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  k_fill_warp<<<ctas, WARPS_PER_CTA * 32, 0, st>>>(p, groups);
  return cudaGetLastError();
}

cudaError_t launch_fill_cta(const FillParams &p, int groups, cudaStream_t st) {
  k_fill_cta<<<groups, CTA_THREADS, 0, st>>>(p, groups);
  return cudaGetLastError();
}

}  // namespace g2s
