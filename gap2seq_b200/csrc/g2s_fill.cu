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

static const int CTA_THREADS = 1024;
// Thin-tier shapes (helper warps for the right pass, main warps for everything else).  A gap's latency falls with more
// warps (the per-level work of a narrow batch is instruction bound on one warp) but fewer gaps are resident then; the
// engine picks per pass: wide shapes for small batches (latency), the two-warp shape for large ones (throughput).
static const int N_THIN_SHAPES = 3;
static const int THIN_MAIN[N_THIN_SHAPES] = {1, 4, 8};
static const int THIN_HELPER[N_THIN_SHAPES] = {1, 2, 4};

struct WarpGroup {
  static constexpr bool kWarp = true;
  static constexpr bool kWarpScout = false;
  static __device__ __forceinline__ int rank() { return threadIdx.x & 31; }
  static __device__ __forceinline__ int size() { return 32; }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
  static __device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
};
struct BlockGroup {
  static constexpr bool kWarp = false;
  static constexpr bool kWarpScout = false;
  static __device__ __forceinline__ int rank() { return threadIdx.x; }
  static __device__ __forceinline__ int size() { return blockDim.x; }
  static __device__ __forceinline__ void sync() { __syncthreads(); }
  static __device__ __forceinline__ bool any(bool p) { return __syncthreads_or(p) != 0; }
};
// NWARPS consecutive warps of the CTA, starting with warp FIRST, synchronised on named barrier BAR.  The per-level
// work of a narrow batch (node table, layout, marking) is instruction bound on one warp, so the thin tier gives each
// pass a few warps; the scout (a chain of dependent gathers) stays on the group's first warp.
template <int BAR, int FIRST, int NWARPS>
struct SubGroup {
  static constexpr bool kWarp = NWARPS == 1;
  static constexpr bool kWarpScout = NWARPS > 1;
  static __device__ __forceinline__ int rank() { return (int)threadIdx.x - 32 * FIRST; }
  static __device__ __forceinline__ int size() { return 32 * NWARPS; }
  static __device__ __forceinline__ void sync() {
    if (NWARPS == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(32 * NWARPS) : "memory");
  }
  static __device__ __forceinline__ bool any(bool p) {
    if (NWARPS == 1) return __any_sync(0xffffffffu, p) != 0;
    unsigned r;
    asm volatile("{\n\t.reg .pred pi, po;\n\tsetp.ne.u32 pi, %1, 0;\n\tbar.red.or.pred po, %2, %3, pi;\n\tselp.u32 %0, 1, 0, po;\n\t}"
                 : "=r"(r) : "r"((unsigned)p), "n"(BAR), "n"(32 * NWARPS) : "memory");
    return r != 0;
  }
};
// named barrier 1 of the thin CTA: the helper warps arrive when the right set is complete, the main warps wait on it
template <int HW, int MW>
struct PairWait {
  static constexpr bool kConcurrent = true;
  static __device__ __forceinline__ void wait_right() {
    SubGroup<2, HW, MW>::sync();
    asm volatile("bar.sync 1, %0;" ::"n"(32 * (HW + MW)) : "memory");
  }
};

template <int HW, int MW>
static __device__ __forceinline__ void process_gap_pair(const FillParams &P, const GapDev &gd, const Slice &Sfull, GShared &sh,
                                                        GapOut *out) {
  typedef SubGroup<3, 0, HW> HelperGroup;     // right pass
  typedef SubGroup<2, HW, MW> MainGroup;      // forward pass, backward marking, pack
  const GapCtx cx = make_ctx(P, gd);
  // The slice's node table is sized for the longest gap of the pass; this gap clears and uses only what its own
  // number of levels calls for (the same rule the engine applies to the pass: 3 entries per level, power of two),
  // which is what keeps the DRAM write traffic of a pass of mostly short gaps near its algorithmic bytes.
  Slice S = Sfull;
  {
    uint32_t want = 3u * ((uint32_t)cx.Dmax + 64u), cap = 256;
    while (cap < want) cap <<= 1;
    if (cap < S.hcap) { S.hcap = cap; S.hmask = cap - 1; }
  }
  unsigned long long t[5];
  t[0] = now_cycles();
  gap_init<BlockGroup>(P, cx, S, sh);
  t[1] = now_cycles();
  if ((threadIdx.x >> 5) < HW) {
    gap_right<HelperGroup>(P, cx, S, sh, sh.nbuf[0]);
    __threadfence_block();
    HelperGroup::sync();
    asm volatile("bar.arrive 1, %0;" ::"n"(32 * (HW + MW)) : "memory");
  } else {
    t[2] = t[1];
    gap_forward<MainGroup, PairWait<HW, MW>>(P, gd, cx, S, sh, sh.nbuf[1]);
    t[3] = now_cycles();
    t[4] = gap_backward_pack<MainGroup>(P, cx, S, sh, sh.nbuf[1]);
    gap_finish<MainGroup>(S, sh, out, t);
  }
}

template <int HW, int MW>   // HW == 0: wide tier, the whole CTA runs the phases one after the other
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
    if (HW > 0) process_gap_pair<(HW > 0 ? HW : 1), MW>(P, gd, S, sh, &P.outs[gi]);
    else process_gap<BlockGroup>(P, gd, S, sh, &P.outs[gi]);
  }
}

template <int HW, int MW>
__global__ void __launch_bounds__(32 * (HW + MW)) k_fill_thin(const __grid_constant__ FillParams P, int groups) {
  __shared__ GShared sh;
  if (blockIdx.x >= (uint32_t)groups) return;
  group_loop<HW, MW>(P, sh, blockIdx.x);
}

__global__ void __launch_bounds__(CTA_THREADS) k_fill_cta(const __grid_constant__ FillParams P, int groups) {
  __shared__ GShared sh;
  if (blockIdx.x >= (uint32_t)groups) return;
  group_loop<0, 1>(P, sh, blockIdx.x);
}

int fill_thin_shapes() { return N_THIN_SHAPES; }
int fill_thin_threads(int shape) { return 32 * (THIN_MAIN[shape] + THIN_HELPER[shape]); }

// occupancy queries cost tens of microseconds each and the answers never change: asked once per shape
int fill_max_resident_thin_groups(int sms, int shape) {
  static int cache[N_THIN_SHAPES] = {0, 0, 0};
  if (shape >= 0 && shape < N_THIN_SHAPES && cache[shape] > 0) return sms * cache[shape];
  int per_sm = 0;
  cudaError_t e = shape == 0   ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fill_thin<1, 1>, 64, 0)
                  : shape == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fill_thin<2, 4>, 192, 0)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fill_thin<4, 8>, 384, 0);
  if (e != cudaSuccess || per_sm < 1) per_sm = 1;
  if (shape >= 0 && shape < N_THIN_SHAPES) cache[shape] = per_sm;
  return sms * per_sm;
}

int fill_max_resident_cta_groups(int sms) {
  static int cached = 0;
  if (cached > 0) return sms * cached;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fill_cta, CTA_THREADS, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  cached = per_sm;
  return sms * per_sm;
}

cudaError_t launch_fill_thin(const FillParams &p, int groups, int shape, cudaStream_t st) {
  if (shape == 0) k_fill_thin<1, 1><<<groups, 64, 0, st>>>(p, groups);
  else if (shape == 1) k_fill_thin<2, 4><<<groups, 192, 0, st>>>(p, groups);
  else k_fill_thin<4, 8><<<groups, 384, 0, st>>>(p, groups);
  return cudaGetLastError();
}

cudaError_t launch_fill_cta(const FillParams &p, int groups, cudaStream_t st) {
  k_fill_cta<<<groups, CTA_THREADS, 0, st>>>(p, groups);
  return cudaGetLastError();
}

}  // namespace g2s
