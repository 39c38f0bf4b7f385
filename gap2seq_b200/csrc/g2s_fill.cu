// g2s_fill.cu - the persistent fill kernels.  Each thread group owns one scratch slice, pulls gaps from a device
// work queue (longest first) and runs the phases of g2s_fill.cuh on them, so many short gaps share one launch.
//   thin tier  k_fill_thin: a CTA of two warps per gap.  After a joint init the helper warp runs the right pass
//              (Gap2Seq.cpp:871-982) while the main warp runs the forward pass (Gap2Seq.cpp:984-1167) up to the first
//              level that consults the right set; a named barrier hands the finished set over.  The main warp then
//              finishes the forward pass, marks backward and packs.
//   wide tier  k_fill_cta: one 512-thread CTA per gap, phases one after the other, for gaps that outgrow the thin
//              tier's scratch slice.
#include <cuda_runtime.h>

#include "g2s_device.h"
#include "g2s_fill.cuh"

namespace g2s {

static const int THIN_THREADS = 64;
static const int CTA_THREADS = 512;

struct WarpGroup {
  static constexpr bool kWarp = true;
  static __device__ __forceinline__ int rank() { return threadIdx.x & 31; }
  static __device__ __forceinline__ int size() { return 32; }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
  static __device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
};
struct BlockGroup {
  static constexpr bool kWarp = false;
  static __device__ __forceinline__ int rank() { return threadIdx.x; }
  static __device__ __forceinline__ int size() { return blockDim.x; }
  static __device__ __forceinline__ void sync() { __syncthreads(); }
  static __device__ __forceinline__ bool any(bool p) { return __syncthreads_or(p) != 0; }
};

// named barrier 1 of the thin CTA: the helper warp arrives when the right set is complete, the main warp waits on it
struct PairWait {
  static constexpr bool kConcurrent = true;
  static __device__ __forceinline__ void wait_right() {
    __syncwarp();
    asm volatile("bar.sync 1, 64;" ::: "memory");
  }
};
static __device__ __forceinline__ void pair_signal_right() {
  __threadfence_block();
  __syncwarp();
  asm volatile("bar.arrive 1, 64;" ::: "memory");
}

static __device__ __forceinline__ void process_gap_pair(const FillParams &P, const GapDev &gd, const Slice &S, GShared &sh,
                                                        GapOut *out) {
  const GapCtx cx = make_ctx(P, gd);
  unsigned long long t[5];
  t[0] = now_cycles();
  gap_init<BlockGroup>(P, cx, S, sh);
  t[1] = now_cycles();
  if ((threadIdx.x >> 5) == 0) {
    gap_right<WarpGroup>(P, cx, S, sh, sh.nbuf[0]);
    pair_signal_right();
  } else {
    t[2] = t[1];
    gap_forward<WarpGroup, PairWait>(P, gd, cx, S, sh, sh.nbuf[1]);
    t[3] = now_cycles();
    t[4] = gap_backward_pack<WarpGroup>(P, cx, S, sh, sh.nbuf[1]);
    gap_finish<WarpGroup>(S, sh, out, t);
  }
}

template <bool PAIR>
__device__ __forceinline__ void group_loop(const FillParams &P, GShared &sh, uint32_t group_id) {
  const Slice S = make_slice(P.scratch + (size_t)group_id * P.lay.bytes, P.lay);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sh.queue_item = atomicAdd(P.queue_head, 1u);
    __syncthreads();
    const uint32_t q = ld_sh(&sh.queue_item);
    if (q >= P.n_gaps) break;
    const uint32_t gi = P.order[q];
    const GapDev gd = P.gaps[gi];
    if (PAIR) process_gap_pair(P, gd, S, sh, &P.outs[gi]);
    else process_gap<BlockGroup>(P, gd, S, sh, &P.outs[gi]);
  }
}

__global__ void __launch_bounds__(THIN_THREADS) k_fill_thin(const __grid_constant__ FillParams P, int groups) {
  __shared__ GShared sh;
  if (blockIdx.x >= (uint32_t)groups) return;
  group_loop<true>(P, sh, blockIdx.x);
}

__global__ void __launch_bounds__(CTA_THREADS) k_fill_cta(const __grid_constant__ FillParams P, int groups) {
  __shared__ GShared sh;
  if (blockIdx.x >= (uint32_t)groups) return;
  group_loop<false>(P, sh, blockIdx.x);
}

int fill_max_resident_thin_groups(int sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fill_thin, THIN_THREADS, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return sms * per_sm;
}

int fill_max_resident_cta_groups(int sms) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fill_cta, CTA_THREADS, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  return sms * per_sm;
}

cudaError_t launch_fill_thin(const FillParams &p, int groups, cudaStream_t st) {
  k_fill_thin<<<groups, THIN_THREADS, 0, st>>>(p, groups);
  return cudaGetLastError();
}

cudaError_t launch_fill_cta(const FillParams &p, int groups, cudaStream_t st) {
  k_fill_cta<<<groups, CTA_THREADS, 0, st>>>(p, groups);
  return cudaGetLastError();
}

}  // namespace g2s
