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
  static __device__ __forceinline__ bool any_word(const uint32_t *w, uint32_t bits) { return any((ld_sh(w) & bits) != 0); }
};
struct BlockGroup {
  static constexpr bool kWarp = false;
  static constexpr bool kWarpScout = false;
  static __device__ __forceinline__ int rank() { return threadIdx.x; }
  static __device__ __forceinline__ int size() { return blockDim.x; }
  static __device__ __forceinline__ void sync() { __syncthreads(); }
  static __device__ __forceinline__ bool any(bool p) { return __syncthreads_or(p) != 0; }
  static __device__ __forceinline__ bool any_word(const uint32_t *w, uint32_t bits) { return any((ld_sh(w) & bits) != 0); }
};
// A thread-block cluster per gap: the wide tier's level step is bound by the scattered global operations one SM keeps in
// flight (profiles/r2_summary.md), and a pass of a few dozen wide gaps leaves most SMs idle, so a gap gets the threads
// of 2 - 8 SMs.  The gap's shared state (GShared) lives in the first CTA's shared memory and the other CTAs reach it
// through distributed shared memory; the group barrier is the cluster barrier (release / acquire at cluster scope, which
// also orders the group's global-memory traffic like __syncthreads does inside a CTA).
struct ClusterGroup {
  static constexpr bool kWarp = false;
  static constexpr bool kWarpScout = true;    // the first warp of the first CTA scouts narrow stretches
  static __device__ __forceinline__ int rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return (int)(r * blockDim.x + threadIdx.x);
  }
  static __device__ __forceinline__ int size() {
    unsigned n;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(n));
    return (int)(n * blockDim.x);
  }
  static __device__ __forceinline__ void sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  // a group decision on a word of the shared state: nobody can change it between the two barriers
  static __device__ __forceinline__ bool any_word(const uint32_t *w, uint32_t bits) {
    sync();
    const bool v = (ld_sh(w) & bits) != 0;
    sync();
    return v;
  }
  static __device__ __forceinline__ bool any(bool p) { return p; }   // not used: decisions go through any_word
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
  static __device__ __forceinline__ bool any_word(const uint32_t *w, uint32_t bits) { return any((ld_sh(w) & bits) != 0); }
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
    if constexpr (HW > 0) process_gap_pair<(HW > 0 ? HW : 1), MW>(P, gd, S, sh, &P.outs[gi]);
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

// wide tier on a cluster of CTAs per gap (cluster size is a launch attribute)
__global__ void __launch_bounds__(CTA_THREADS) k_fill_cluster(const __grid_constant__ FillParams P, int groups) {
  __shared__ GShared sh_local;
  unsigned ncta, crank;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(ncta));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const uint32_t group_id = blockIdx.x / ncta;
  // the first CTA's copy is the gap's shared state
  GShared *shp;
  {
    unsigned long long local = (unsigned long long)__cvta_generic_to_shared(&sh_local), remote;
    asm volatile("mapa.shared::cluster.u64 %0, %1, %2;" : "=l"(remote) : "l"(local), "r"(0u));
    unsigned long long gen;
    asm volatile("cvta.shared::cluster.u64 %0, %1;" : "=l"(gen) : "l"(remote));
    shp = (GShared *)gen;
  }
  GShared &sh = *shp;
  if (group_id < (uint32_t)groups) {
    const Slice S = make_slice(P.scratch + (size_t)group_id * P.lay.bytes, P.lay);
    for (;;) {
      ClusterGroup::sync();
      if (crank == 0 && threadIdx.x == 0) sh.queue_item = atomicAdd(P.queue_head, 1u);
      ClusterGroup::sync();
      const uint32_t q = ld_sh(&sh.queue_item);
      if (q >= P.n_gaps) break;
      const uint32_t gi = P.order[q];
      const GapDev gd = P.gaps[gi];
      process_gap<ClusterGroup>(P, gd, S, sh, &P.outs[gi]);
    }
  }
  ClusterGroup::sync();   // no CTA leaves while another may still touch its shared memory
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

// clusters of `cluster` CTAs that can be resident at once (0: this cluster size cannot be launched)
int fill_max_resident_clusters(int cluster) {
  static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cluster < 2 || cluster > 8) return 0;
  if (cache[cluster] != 0) return cache[cluster] > 0 ? cache[cluster] : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cluster * 64, 1, 1);
  cfg.blockDim = dim3(CTA_THREADS, 1, 1);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_fill_cluster, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[cluster] = n > 0 ? n : -1;
  return n > 0 ? n : 0;
}

cudaError_t launch_fill_cluster(const FillParams &p, int groups, int cluster, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * cluster), 1, 1);
  cfg.blockDim = dim3(CTA_THREADS, 1, 1);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_fill_cluster, p, groups);
}

cudaError_t launch_fill_cta(const FillParams &p, int groups, cudaStream_t st) {
  k_fill_cta<<<groups, CTA_THREADS, 0, st>>>(p, groups);
  return cudaGetLastError();
}

}  // namespace g2s
