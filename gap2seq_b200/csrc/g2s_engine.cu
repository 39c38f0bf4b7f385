// g2s_engine.cu - host side of libg2s_gpu.so: the C ABI of include/g2s.h, batching of gaps into device passes
// (thin tier - a two-warp CTA per gap -, then wide CTA-per-gap tiers for gaps that outgrow their scratch slice), and the two
// stages of Gap2Seq::fill_gap that are inherently sequential on the host:
//   * the SCC contraction / branch counter of src/Gap2Seq.cpp:1314-1434 for the (rare) gaps whose path subgraph
//     is not a plain level DAG (the device classifies the DAG case itself),
//   * the random traceback of src/Gap2Seq.cpp:1437-1517, which must consume libc rand() in input order
//      and therefore walks the device-built predecessor trellis gap after gap.
// There is no CPU implementation of the DP here: without a CUDA device every call fails with G2S_E_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <malloc.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/g2s.h"
#include "g2s_device.h"
#include "g2s_fill.cuh"
#include "g2s_chain.cuh"
#include "g2s_finish.h"
#include "g2s_h5.h"
#include "g2s_scaffold.h"

using namespace g2s;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                                         \
  do {                                                                                                      \
    cudaError_t e_ = (x);                                                                                   \
    if (e_ != cudaSuccess)                                                                                  \
      return fail(e_ == cudaErrorMemoryAllocation ? G2S_E_NOMEM : G2S_E_CUDA,                               \
                  std::string(#x) + ": " + cudaGetErrorString(e_));                                         \
  } while (0)

namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Pass {                   // one device pass over a subset of the batch
  int tier = 0;
  int shape = 0;                // thin tier: warps per gap (g2s_device.h)
  int cluster = 1;              // wide tiers: CTAs (SMs) per gap
  std::vector<uint32_t> order;  // gap ids, longest first
  SliceLayout lay{};
  int groups = 0;
  unsigned long long pack_cap = 0;
  unsigned long long pack_used = 0;
  int pack_slot = 0;            // index into the engine's persistent device pack buffers
  unsigned long long fill_cap = 0, fill_used = 0;   // fill arena of the device-side tracebacks (bytes)
};

struct PinPool {                // pinned host staging for pack arenas, shared by the engine and its detached batches
  std::mutex mu;
  std::vector<PinBuf> free;
  ~PinPool() { for (auto &b : free) b.release(); }
};

}  // namespace

struct g2s_engine {
  g2s_config cfg{};
  int sms = 0;
  size_t budget = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[4]{};
  DeviceTable table;
  // batch state
  std::vector<g2s_gap> gaps;           // pointers rewired into flank_store
  std::vector<char> flank_store;
  std::vector<GapDev> gapdev;
  std::vector<GapOut> outs;
  std::vector<int> out_pass;           // pass that produced outs[i]
  std::vector<uint8_t> phantom;        // gaps whose g2s_gap.present_bits disagree with the table (empty: no gap carried bits)
  std::vector<Pass> passes;
  bool uploaded = false, ran = false;
  DevBuf d_flanks, d_gaps, d_order, d_outs, d_scratch, d_counters;
  MultiTable multi;                    // per-gap filtered-reads mode: one graph per gap group (g2s_build_graphs_from_reads)
  DevBuf d_graph_of;                   // graph of gap i of the uploaded batch
  DevBuf d_chain_strs, d_chain_walks;  // chain tier: packed walk strings, walk summaries
  bool chain_tier = true;              // G2S_NO_CHAIN=1: every gap starts in the thin tier (A/B measurements)
  std::vector<DevBuf> d_pack;          // persistent (grow-only) pack arena per pass ordinal
  std::vector<DevBuf> d_fill;          // fill arena per pass ordinal
  bool device_trace = true;            // G2S_NO_DEVICE_TRACE=1: every traceback on the host (A/B measurements)
  std::shared_ptr<PinPool> pin = std::make_shared<PinPool>();
  PinBuf h_outs, h_stage;
  g2s_timing tm{};
  int max_dmax = 0, max_rmf = 0;
  RandAhead ahead;                     // persistent host buffers of the traceback stage
  RawBuf<char> trace_arena;
  std::unique_ptr<ThreadPool> pool;    // host threads (g2s_config.host_threads) for the stages after a device pass
  // g2s_create_multi: the engines of the other GPUs of the box.  This engine (devices[0]) owns the graph import, the
  // sharding of every wave of gaps, the single libc rand() stream of the tracebacks and the output; a peer runs the
  // device pass + D2H + subgraph analysis of its shard on its own host thread.
  std::vector<g2s_engine *> peers;
  std::string table_transport;         // how the table reached the peers ("nccl" / "peer-copy"), for the log
  double table_replicate_ms = 0;
};

// ------------------------------------------------------------------------------------------------ C ABI: basics
extern "C" const char *g2s_last_error(void) { return g_err.c_str(); }
extern "C" int g2s_abi_version(void) { return G2S_ABI_VERSION; }

extern "C" int g2s_create(const g2s_config *cfg, g2s_engine **out) {
  if (!cfg || !out) return fail(G2S_E_ARG, "g2s_create: null argument");
  if (cfg->abi_version != G2S_ABI_VERSION) return fail(G2S_E_ARG, "g2s_create: ABI version mismatch");
  if (cfg->k < 3 || cfg->k > 31 || !(cfg->k & 1))
    return fail(G2S_E_ARG, "g2s_create: k must be odd and in [3, 31] (even k has palindromic k-mers, SURVEY Q11)");
  if (cfg->max_fuz < 0 || cfg->max_fuz > MAX_FUZ) return fail(G2S_E_ARG, "g2s_create: max_fuz must be in [0, 31]");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(G2S_E_CUDA, "g2s_create: no such CUDA device");
  CUDA_TRY(cudaSetDevice(cfg->device));
  g2s_engine *e = new g2s_engine();
  e->cfg = *cfg;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  e->sms = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
  for (auto &v : e->ev) CUDA_TRY(cudaEventCreate(&v));
  size_t fr = 0, tot = 0;
  CUDA_TRY(cudaMemGetInfo(&fr, &tot));
  e->budget = cfg->max_mem_bytes > 0 ? (size_t)cfg->max_mem_bytes : (size_t)(fr * 0.6);
  CUDA_TRY(e->d_counters.reserve(64));
  e->pool.reset(new ThreadPool(cfg->host_threads));
  // The driver call hands megabytes of malloc'ed strings to the caller every time; with glibc's default trimming the
  // freed memory goes back to the kernel and every call pays its page faults again (several ms per 10 000 records on
  // the measured host).  Keeping freed memory in the process is a process-wide policy and therefore the APPLICATION's
  // decision: Gap2Seq-core (g2s_core_main.cpp) and bench.py make it for themselves; a library user asks for it with
  // G2S_TUNE_MALLOC=1.
  if (getenv("G2S_TUNE_MALLOC")) { mallopt(M_TRIM_THRESHOLD, 1 << 30); mallopt(M_MMAP_THRESHOLD, 1 << 28); }
  e->device_trace = getenv("G2S_NO_DEVICE_TRACE") == nullptr;
  e->chain_tier = e->device_trace && getenv("G2S_NO_CHAIN") == nullptr;   // the chain tier writes fills on the device
  *out = e;
  return G2S_OK;
}

extern "C" void g2s_destroy(g2s_engine *e) {
  if (!e) return;
  for (g2s_engine *p : e->peers) g2s_destroy(p);
  e->peers.clear();
  cudaSetDevice(e->cfg.device);
  e->d_chain_strs.release(); e->d_chain_walks.release();
  for (auto &b : e->d_pack) b.release();
  for (auto &b : e->d_fill) b.release();
  e->d_flanks.release(); e->d_gaps.release(); e->d_order.release(); e->d_outs.release();
  e->d_scratch.release(); e->d_counters.release(); e->h_outs.release(); e->h_stage.release();
  if (e->table.slots) cudaFree(e->table.slots);
  multi_table_release(&e->multi); e->d_graph_of.release();
  for (auto &v : e->ev) if (v) cudaEventDestroy(v);
  if (e->st) cudaStreamDestroy(e->st);
  delete e;
}


// ------------------------------------------------------------------------------------------------ several GPUs
// The k-mer table is built (or imported) on devices[0] and replicated to the peers ONCE, by one ncclBroadcast over
// NVLink / NVSwitch.  NCCL is bound at run time (dlopen of libnccl.so.2: the library the process already
// loaded - e.g. torch's - or the system one), so that libg2s_gpu.so itself has no link-time dependency; where no NCCL
// library can be loaded the table travels by cudaMemcpyPeerAsync over the same links.
namespace {
struct NcclApi {
  typedef struct ncclComm *comm_t;
  int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int /*ncclDataType_t*/, int, comm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
  NcclApi() {
    if (getenv("G2S_NO_NCCL")) return;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
    Broadcast = (decltype(Broadcast))dlsym(h, "ncclBroadcast");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    ok = CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast && GetErrorString;
  }
};
NcclApi &nccl() { static NcclApi a; return a; }
}  // namespace

static int replicate_table(g2s_engine *e) {
  if (e->peers.empty() || !e->table.slots) return G2S_OK;
  const auto t0 = std::chrono::steady_clock::now();
  const size_t bytes = e->table.cap * sizeof(Slot);
  for (g2s_engine *p : e->peers) {
    CUDA_TRY(cudaSetDevice(p->cfg.device));
    if (p->table.slots) { cudaFree(p->table.slots); p->table.slots = nullptr; }
    CUDA_TRY(cudaMalloc(&p->table.slots, bytes));
    p->table.cap = e->table.cap; p->table.n = e->table.n; p->table.k = e->table.k;
  }
  const int n = 1 + (int)e->peers.size();
  bool done = false, dup = false;
  for (g2s_engine *p : e->peers) {
    if (p->cfg.device == e->cfg.device) dup = true;
    for (g2s_engine *q : e->peers) if (q != p && q->cfg.device == p->cfg.device) dup = true;
  }
  if (!dup && nccl().ok) {
    std::vector<int> devs(n);
    std::vector<g2s_engine *> all(n);
    all[0] = e; devs[0] = e->cfg.device;
    for (int i = 1; i < n; i++) { all[i] = e->peers[i - 1]; devs[i] = all[i]->cfg.device; }
    std::vector<NcclApi::comm_t> comms(n, nullptr);
    int rc = nccl().CommInitAll(comms.data(), n, devs.data());
    if (rc == 0) {
      rc = nccl().GroupStart();
      for (int i = 0; i < n && rc == 0; i++) {
        cudaSetDevice(devs[i]);
        rc = nccl().Broadcast(e->table.slots, all[i]->table.slots, bytes, 0 /* ncclChar */, 0, comms[i], all[i]->st);
      }
      const int rc2 = nccl().GroupEnd();
      if (rc == 0) rc = rc2;
      for (int i = 0; i < n; i++) { cudaSetDevice(devs[i]); cudaError_t ce = cudaStreamSynchronize(all[i]->st); if (ce != cudaSuccess && rc == 0) rc = -1; }
      for (int i = 0; i < n; i++) if (comms[i]) nccl().CommDestroy(comms[i]);
    }
    if (rc != 0) return fail(G2S_E_CUDA, std::string("table replication (NCCL broadcast): ") + (rc > 0 ? nccl().GetErrorString(rc) : "CUDA error"));
    done = true;
    e->table_transport = "nccl";
  }
  if (!done) {
    for (g2s_engine *p : e->peers) {
      CUDA_TRY(cudaSetDevice(e->cfg.device));
      int can = 0;
      cudaDeviceCanAccessPeer(&can, e->cfg.device, p->cfg.device);
      if (can) { cudaError_t ce = cudaDeviceEnablePeerAccess(p->cfg.device, 0); if (ce != cudaSuccess) cudaGetLastError(); }
      CUDA_TRY(cudaMemcpyPeerAsync(p->table.slots, p->cfg.device, e->table.slots, e->cfg.device, bytes, e->st));
    }
    CUDA_TRY(cudaStreamSynchronize(e->st));
    e->table_transport = "peer-copy";
  }
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  e->table_replicate_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (getenv("G2S_HOST_PROF"))
    fprintf(stderr, "g2s: table of %.1f MB replicated to %zu peer GPUs by %s in %.1f ms\n", bytes / 1e6, e->peers.size(),
            e->table_transport.c_str(), e->table_replicate_ms);
  return G2S_OK;
}

extern "C" int g2s_create_multi(const g2s_config *cfg, const int32_t *devices, int32_t n_devices, g2s_engine **out) {
  if (!cfg || !out || !devices || n_devices < 1) return fail(G2S_E_ARG, "g2s_create_multi: bad argument");
  // G2S_ALLOW_DUPLICATE_DEVICES=1 (tests on a one-GPU box): several engines on one GPU exercise the sharding, the
  // gather in input order and the table replication (device-to-device copy; NCCL refuses duplicate devices)
  bool dup = false;
  for (int i = 0; i < n_devices; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) dup = true;
  if (dup && !getenv("G2S_ALLOW_DUPLICATE_DEVICES")) return fail(G2S_E_ARG, "g2s_create_multi: device listed twice");
  g2s_config c0 = *cfg;
  c0.device = devices[0];
  g2s_engine *e = nullptr;
  int rc = g2s_create(&c0, &e);
  if (rc != G2S_OK) return rc;
  // the host threads are shared out: the primary keeps its pool for the tracebacks and the splice of the whole batch,
  // every peer gets a share for the subgraph analysis of its shard
  const int total = e->pool->size();
  for (int i = 1; i < n_devices; i++) {
    g2s_config ci = *cfg;
    ci.device = devices[i];
    ci.host_threads = std::max(1, total / n_devices);
    g2s_engine *p = nullptr;
    rc = g2s_create(&ci, &p);
    if (rc != G2S_OK) { g2s_destroy(e); return rc; }
    e->peers.push_back(p);
  }
  cudaSetDevice(e->cfg.device);
  *out = e;
  return G2S_OK;
}

extern "C" int g2s_num_devices(const g2s_engine *e) { return e ? 1 + (int)e->peers.size() : -1; }

extern "C" int g2s_import_kmers(g2s_engine *e, const uint64_t *canon, size_t n) {
  if (!e || (!canon && n)) return fail(G2S_E_ARG, "g2s_import_kmers: null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  uint64_t *d_keys = nullptr;
  multi_table_release(&e->multi);
  CUDA_TRY(table_alloc(&e->table, e->cfg.k, n, e->sms, e->st));
  if (n) {
    CUDA_TRY(cudaMalloc(&d_keys, n * 8));
    cudaError_t er = cudaMemcpyAsync(d_keys, canon, n * 8, cudaMemcpyHostToDevice, e->st);
    if (er == cudaSuccess) er = table_fill(&e->table, d_keys, nullptr, 0, n, e->sms, e->st);
    cudaFree(d_keys);
    CUDA_TRY(er);
  } else {
    CUDA_TRY(cudaStreamSynchronize(e->st));
  }
  return replicate_table(e);
}

extern "C" int g2s_h5_read_solid(const char *path, uint64_t **canon_out, size_t *n_out, int32_t *kmer_size_out) {
  if (!path || !canon_out || !n_out) return fail(G2S_E_ARG, "g2s_h5_read_solid: null argument");
  *canon_out = nullptr; *n_out = 0;
  FILE *f = fopen(path, "rb");
  if (!f) return fail(G2S_E_ARG, std::string("g2s_h5_read_solid: cannot open ") + path);
  std::vector<uint8_t> data;
  {
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    data.resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = data.empty() ? 0 : fread(data.data(), 1, data.size(), f);
    fclose(f);
    if (got != data.size()) return fail(G2S_E_ARG, std::string("g2s_h5_read_solid: short read of ") + path);
  }
  std::vector<uint64_t> kmers;
  int ks = 0;
  try { h5::read_solid_kmers(data.data(), data.size(), kmers, nullptr, &ks); }
  catch (const std::exception &ex) { return fail(G2S_E_ARG, std::string(ex.what()) + " (" + path + ")"); }
  if (kmer_size_out) *kmer_size_out = ks;
  uint64_t *out = (uint64_t *)malloc(std::max<size_t>(kmers.size(), 1) * 8);
  if (!out) return fail(G2S_E_ARG, "g2s_h5_read_solid: out of memory");
  if (!kmers.empty()) memcpy(out, kmers.data(), kmers.size() * 8);
  *canon_out = out; *n_out = kmers.size();
  return G2S_OK;
}

extern "C" int g2s_import_h5(g2s_engine *e, const char *path) {
  if (!e) return fail(G2S_E_ARG, "g2s_import_h5: null argument");
  uint64_t *canon = nullptr; size_t n = 0; int32_t ks = 0;
  const int rc = g2s_h5_read_solid(path, &canon, &n, &ks);
  if (rc != G2S_OK) return rc;
  int out = G2S_OK;
  if (ks != 0 && ks != e->cfg.k) out = fail(G2S_E_ARG, "g2s_import_h5: the graph file holds " + std::to_string(ks) + "-mers, the engine was created for k = " + std::to_string(e->cfg.k));
  else out = g2s_import_kmers(e, canon, n);
  free(canon);
  return out;
}

extern "C" int g2s_build_from_reads(g2s_engine *e, const char *bases, const uint64_t *offsets, size_t n_reads, int solid) {
  if (!e || !bases || !offsets) return fail(G2S_E_ARG, "g2s_build_from_reads: null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  multi_table_release(&e->multi);
  CUDA_TRY(table_from_reads(&e->table, e->cfg.k, bases, offsets, n_reads, solid, e->sms, e->st));
  return replicate_table(e);
}

extern "C" int g2s_build_from_read_files(g2s_engine *e, const char *const *paths, size_t n_paths, int solid, uint64_t *n_reads_out) {
  if (!e || !paths || n_paths == 0) return fail(G2S_E_ARG, "g2s_build_from_read_files: null argument");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  // the files as they are, one after the other (each made to end with a line break); the device does the parsing
  std::vector<char> bytes;
  for (size_t f = 0; f < n_paths; f++) {
    FILE *fp = paths[f] ? fopen(paths[f], "rb") : nullptr;
    if (!fp) return fail(G2S_E_ARG, std::string("g2s_build_from_read_files: cannot open ") + (paths[f] ? paths[f] : "(null)"));
    fseek(fp, 0, SEEK_END);
    const long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    const size_t at = bytes.size();
    bytes.resize(at + (sz > 0 ? (size_t)sz : 0));
    const size_t got = sz > 0 ? fread(bytes.data() + at, 1, (size_t)sz, fp) : 0;
    fclose(fp);
    if (got != (size_t)(sz > 0 ? sz : 0)) return fail(G2S_E_ARG, std::string("g2s_build_from_read_files: short read of ") + paths[f]);
    if (!bytes.empty() && bytes.back() != '\n') bytes.push_back('\n');
  }
  multi_table_release(&e->multi);
  uint64_t nr = 0;
  const cudaError_t er = table_from_read_files(&e->table, e->cfg.k, bytes.data(), bytes.size(), solid, e->sms, e->st, &nr);
  if (er == cudaErrorInvalidValue) return fail(G2S_E_ARG, "g2s_build_from_read_files: the reads are neither FASTA ('>') nor FASTQ ('@') text");
  CUDA_TRY(er);
  if (n_reads_out) *n_reads_out = nr;
  return replicate_table(e);
}

extern "C" int g2s_build_graphs_from_reads(g2s_engine *e, const char *bases, const uint64_t *offsets, size_t n_reads,
                                           const uint64_t *graph_first_read, size_t n_graphs, int solid) {
  if (!e || !offsets || !graph_first_read || (!bases && n_reads)) return fail(G2S_E_ARG, "g2s_build_graphs_from_reads: null argument");
  if (!e->peers.empty()) return fail(G2S_E_STATE, "g2s_build_graphs_from_reads: per-gap graphs are built on a one-GPU engine");
  if (n_graphs == 0) return fail(G2S_E_ARG, "g2s_build_graphs_from_reads: no graph");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  if (e->table.slots) { cudaFree(e->table.slots); e->table = DeviceTable{}; }   // the engine holds one kind of graph at a time
  const cudaError_t er = tables_from_reads(&e->multi, e->cfg.k, bases, offsets, n_reads, graph_first_read, n_graphs, solid, e->sms, e->st);
  if (er == cudaErrorInvalidValue) return fail(G2S_E_ARG, "g2s_build_graphs_from_reads: read ranges must tile the reads in order, fewer than 2^31 k-mer occurrences per call");
  CUDA_TRY(er);
  e->uploaded = false; e->ran = false;
  return G2S_OK;
}

extern "C" int64_t g2s_graph_num_kmers(const g2s_engine *e, size_t graph) {
  if (!e || graph >= e->multi.h_sizes.size()) return -1;
  return (int64_t)e->multi.h_sizes[graph];
}

extern "C" int g2s_table_raw(g2s_engine *e, void **device_ptr, uint64_t *capacity_slots, uint64_t *n_kmers, uint64_t *bytes) {
  if (!e || !device_ptr) return fail(G2S_E_ARG, "g2s_table_raw: null argument");
  if (!e->table.slots) return fail(G2S_E_STATE, "g2s_table_raw: no graph imported");
  *device_ptr = e->table.slots;
  if (capacity_slots) *capacity_slots = e->table.cap;
  if (n_kmers) *n_kmers = e->table.n;
  if (bytes) *bytes = e->table.cap * sizeof(Slot);
  return G2S_OK;
}

extern "C" int g2s_table_adopt(g2s_engine *e, uint64_t capacity_slots, uint64_t n_kmers, void **device_ptr) {
  if (!e || !device_ptr) return fail(G2S_E_ARG, "g2s_table_adopt: null argument");
  if (capacity_slots < 2 || (capacity_slots & (capacity_slots - 1))) return fail(G2S_E_ARG, "g2s_table_adopt: capacity must be a power of two");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  if (e->table.slots) { cudaFree(e->table.slots); e->table.slots = nullptr; }
  CUDA_TRY(cudaMalloc(&e->table.slots, capacity_slots * sizeof(Slot)));
  e->table.cap = capacity_slots; e->table.n = n_kmers; e->table.k = e->cfg.k;
  *device_ptr = e->table.slots;
  return G2S_OK;
}

extern "C" int64_t g2s_num_kmers(const g2s_engine *e) { return e ? (int64_t)e->table.n : -1; }

extern "C" int g2s_export_kmers(g2s_engine *e, uint64_t *canon_sorted, uint8_t *adjacency) {
  if (!e || !canon_sorted) return fail(G2S_E_ARG, "g2s_export_kmers: null argument");
  if (!e->table.slots) return fail(G2S_E_STATE, "g2s_export_kmers: no graph imported");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  CUDA_TRY(table_export(e->table, canon_sorted, adjacency, e->sms, e->st));
  return G2S_OK;
}

extern "C" int g2s_contains(g2s_engine *e, const char *kmers_ascii, size_t n, uint8_t *out) {
  if (!e || !kmers_ascii || !out) return fail(G2S_E_ARG, "g2s_contains: null argument");
  if (!e->table.slots) return fail(G2S_E_STATE, "g2s_contains: no graph imported");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  CUDA_TRY(table_contains(e->table, kmers_ascii, n, out, e->st));
  return G2S_OK;
}

extern "C" int g2s_get_timing(const g2s_engine *e, g2s_timing *out) {
  if (!e || !out) return fail(G2S_E_ARG, "g2s_get_timing: null argument");
  *out = e->tm;
  return G2S_OK;
}

extern "C" int g2s_get_phase_cycles(const g2s_engine *e, double *sum, double *slowest) {
  if (!e || !sum || !slowest) return fail(G2S_E_ARG, "g2s_get_phase_cycles: null argument");
  for (int q = 0; q < 14; q++) sum[q] = slowest[q] = 0;
  uint32_t worst = 0;
  for (const GapOut &o : e->outs) {
    for (int q = 0; q < 6; q++) sum[q] += 16.0 * o.cyc[q];
    for (int q = 0; q < 8; q++) sum[6 + q] += (q < 5 ? 16.0 : 1.0) * o.prof[q];
    if (o.cyc[5] >= worst) {
      worst = o.cyc[5];
      for (int q = 0; q < 6; q++) slowest[q] = 16.0 * o.cyc[q];
      for (int q = 0; q < 8; q++) slowest[6 + q] = (q < 5 ? 16.0 : 1.0) * o.prof[q];
    }
  }
  return G2S_OK;
}

// ------------------------------------------------------------------------------------------------ device passes
static int dmax_of(const g2s_gap &g) { return g.gap_len + g.gap_err + g.left_max_fuz + g.right_max_fuz; }

static int plan_pass(g2s_engine *e, Pass &p) {
  // scratch geometry of this tier
  uint64_t cap0 = (uint64_t)4 * ((uint64_t)e->max_dmax + 2) + 1024;
  if (cap0 < 4096) cap0 = 4096;
  uint64_t cap = cap0;
  for (int t = 0; t < p.tier; t++) cap *= 16;
  uint32_t hlog = p.tier == 0 ? ceil_log2(3 * ((uint64_t)e->max_dmax + 64)) : std::max<uint32_t>(16, ceil_log2(cap / 4));
  p.lay = make_layout(cap, hlog, (uint32_t)e->max_dmax + 1, (uint32_t)e->max_rmf + 1);
  if (cap >= (1ULL << 31) || p.lay.bytes > e->budget) return -1;   // largest tier exceeded
  size_t fit = e->budget / p.lay.bytes;
  int resident = fill_max_resident_cta_groups(e->sms);
  if (p.tier == 0) {
    // Thin tier: pick the CTA shape.  A gap's latency is about levels x cycles-per-level of the shape (measured on
    // B200, cfg4: 870 / 560 / 480 cycles for 2 / 6 / 12 warps); the pass takes the longest gap's latency or, when more
    // gaps than resident CTAs exist, the total work spread over the resident CTAs - whichever is larger.
    static const double cyc_per_level[3] = {870.0, 560.0, 480.0};
    double sum_levels = 0, max_levels = 0;
    for (uint32_t gi : p.order) { const double d = dmax_of(e->gaps[gi]) + 64.0; sum_levels += d; max_levels = std::max(max_levels, d); }
    const char *force = getenv("G2S_THIN_SHAPE");
    double best = 0;
    for (int sh = 0; sh < fill_thin_shapes(); sh++) {
      const int r = fill_max_resident_thin_groups(e->sms, sh);
      const double slots = (double)std::max<size_t>(1, std::min<size_t>({p.order.size(), (size_t)r, fit}));
      const double t = std::max(max_levels, sum_levels / slots) * cyc_per_level[sh];
      if ((force && atoi(force) == sh) || (!force && (sh == 0 || t < best))) { best = t; p.shape = sh; resident = r; }
    }
  }
  p.cluster = 1;
  if (p.tier > 0) {
    // Wide tiers: a pass of few gaps leaves SMs idle while every gap is bound by what ONE SM keeps in flight - give each
    // gap the largest cluster that still lets all gaps of the pass run at once (G2S_CLUSTER forces a size, 1 = none).
    const size_t n_run = std::min<size_t>(p.order.size(), fit);
    const char *force = getenv("G2S_CLUSTER");
    for (int c : {8, 4, 2}) {
      if (force ? atoi(force) != c : (size_t)fill_max_resident_clusters(c) < n_run) continue;
      if (fill_max_resident_clusters(c) < 1) continue;
      p.cluster = c;
      resident = fill_max_resident_clusters(c);
      break;
    }
  }
  size_t groups = std::min<size_t>({p.order.size(), (size_t)resident, fit});
  p.groups = (int)std::max<size_t>(groups, 1);
  // output arena
  unsigned long long want = 0;
  if (p.tier == 0) {
    for (uint32_t gi : p.order) want += 2ULL * ((unsigned long long)dmax_of(e->gaps[gi]) + 256);
  } else {
    want = (unsigned long long)p.order.size() * cap;
  }
  unsigned long long lim = (unsigned long long)(e->budget / 4 / sizeof(PackRec));
  p.pack_cap = std::max<unsigned long long>(std::min(want, lim), 1 << 16);
  // diagnostic knob (tests): a deliberately small pack arena, so that the GS_OUT_FULL -> next tier path can be driven
  if (const char *pc = getenv("G2S_PACK_CAP_RECORDS")) p.pack_cap = std::max<unsigned long long>(1, strtoull(pc, nullptr, 10));
  return 0;
}

static int run_pass(g2s_engine *e, Pass &p) {
  CUDA_TRY(e->d_scratch.reserve((size_t)p.groups * p.lay.bytes));
  if ((size_t)p.pack_slot >= e->d_pack.size()) e->d_pack.resize(p.pack_slot + 1);
  if (e->d_pack[p.pack_slot].cap < p.pack_cap * sizeof(PackRec)) CUDA_TRY(e->d_pack[p.pack_slot].reserve(p.pack_cap * sizeof(PackRec)));
  else if (!getenv("G2S_PACK_CAP_RECORDS")) p.pack_cap = e->d_pack[p.pack_slot].cap / sizeof(PackRec);   // use all of a buffer that is already there
  if ((size_t)p.pack_slot >= e->d_fill.size()) e->d_fill.resize(p.pack_slot + 1);
  p.fill_cap = 0;
  if (e->device_trace) {
    for (uint32_t gi : p.order) p.fill_cap += (unsigned long long)dmax_of(e->gaps[gi]) + 2;
    CUDA_TRY(e->d_fill[p.pack_slot].reserve((size_t)p.fill_cap + 16));
  }
  CUDA_TRY(e->d_order.reserve(p.order.size() * 4));
  CUDA_TRY(cudaMemcpyAsync(e->d_order.p, p.order.data(), p.order.size() * 4, cudaMemcpyHostToDevice, e->st));
  CUDA_TRY(cudaMemsetAsync(e->d_counters.p, 0, 64, e->st));
  FillParams P{};
  P.table = table_view(e->table);
  if (e->multi.d_tables) { P.table = e->multi.h_tables[0]; P.tables = e->multi.d_tables; P.graph_of = (const uint32_t *)e->d_graph_of.p; }
  P.flanks = (const char *)e->d_flanks.p;
  P.gaps = (const GapDev *)e->d_gaps.p;
  P.order = (const uint32_t *)e->d_order.p;
  P.n_gaps = (uint32_t)p.order.size();
  P.outs = (GapOut *)e->d_outs.p;
  P.scratch = (unsigned char *)e->d_scratch.p;
  P.lay = p.lay;
  P.pack = (PackRec *)e->d_pack[p.pack_slot].p;
  P.pack_cap = p.pack_cap;
  P.pack_cursor = (unsigned long long *)((char *)e->d_counters.p + 8);
  P.queue_head = (uint32_t *)e->d_counters.p;
  P.fills = e->device_trace ? (char *)e->d_fill[p.pack_slot].p : nullptr;
  P.fill_cap = p.fill_cap;
  P.fill_cursor = (unsigned long long *)((char *)e->d_counters.p + 16);
  CUDA_TRY(cudaEventRecord(e->ev[0], e->st));
  if (p.tier == 0) CUDA_TRY(launch_fill_thin(P, p.groups, p.shape, e->st));
  else if (p.cluster > 1) CUDA_TRY(launch_fill_cluster(P, p.groups, p.cluster, e->st));
  else CUDA_TRY(launch_fill_cta(P, p.groups, e->st));
  CUDA_TRY(cudaEventRecord(e->ev[1], e->st));
  e->tm.launches++;
  // the per-gap summaries decide whether another tier is needed
  CUDA_TRY(cudaMemcpyAsync(e->h_outs.p, e->d_outs.p, e->gaps.size() * sizeof(GapOut), cudaMemcpyDeviceToHost, e->st));
  unsigned long long used = 0, fused = 0;
  CUDA_TRY(cudaMemcpyAsync(&used, (char *)e->d_counters.p + 8, 8, cudaMemcpyDeviceToHost, e->st));
  CUDA_TRY(cudaMemcpyAsync(&fused, (char *)e->d_counters.p + 16, 8, cudaMemcpyDeviceToHost, e->st));
  CUDA_TRY(cudaStreamSynchronize(e->st));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
  e->tm.kernel_ms += ms;
  if (p.tier == 0) { e->tm.kernel_ms_thin += ms; e->tm.launches_thin++; e->tm.thin_shape = (uint32_t)p.shape; }
  else { e->tm.kernel_ms_wide += ms; e->tm.launches_wide++; e->tm.wide_cluster = (uint32_t)p.cluster; }
  if (getenv("G2S_HOST_PROF"))
    fprintf(stderr, "g2s: pass tier %d (shape %d, cluster %d): %zu gaps on %d groups, slice %.1f MB, %.3f ms\n", p.tier, p.shape,
            p.cluster, p.order.size(), p.groups, p.lay.bytes / 1e6, ms);
  e->tm.d2h_bytes += e->gaps.size() * sizeof(GapOut) + 8;
  p.pack_used = std::min(used, p.pack_cap);
  p.fill_used = std::min(fused, p.fill_cap);
  return G2S_OK;
}

// The chain tier (g2s_chain.cuh) over the gaps of p.order: two launches per chunk of gaps (walks, finish).  The walk
// strings are the tier's only scratch: 2 bits per level and pass, a few hundred bytes per gap.
static int run_chain_pass(g2s_engine *e, Pass &p) {
  const int k = e->cfg.k;
  ChainParams P{};
  P.table = table_view(e->table);
  if (e->multi.d_tables) { P.table = e->multi.h_tables[0]; P.tables = e->multi.d_tables; P.graph_of = (const uint32_t *)e->d_graph_of.p; }
  P.walk_stride = 8;
  if (const char *ws = getenv("G2S_WALK_STRIDE")) { const int v = atoi(ws); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) P.walk_stride = (uint32_t)v; }
  P.f_words = chain_words(k, e->max_dmax);
  P.r_words = chain_words(k, e->max_rmf + (e->max_dmax + 1) / 2);   // right_half <= rmf + ceil((g + e) / 2)
  const size_t stride = ((size_t)P.f_words + P.r_words) * 8;
  size_t chunk = std::max<size_t>(1, std::min<size_t>(p.order.size(), (e->budget / 2) / stride));
  CUDA_TRY(e->d_chain_strs.reserve(chunk * stride));
  CUDA_TRY(e->d_chain_walks.reserve(chunk * 2 * sizeof(ChainWalk)));
  if ((size_t)p.pack_slot >= e->d_pack.size()) e->d_pack.resize(p.pack_slot + 1);
  if ((size_t)p.pack_slot >= e->d_fill.size()) e->d_fill.resize(p.pack_slot + 1);
  p.fill_cap = 0;
  for (uint32_t gi : p.order) p.fill_cap += (unsigned long long)dmax_of(e->gaps[gi]) + 16;   // the tier allocates fills in multiples of 8 bytes
  CUDA_TRY(e->d_fill[p.pack_slot].reserve((size_t)p.fill_cap + 16));
  CUDA_TRY(e->d_order.reserve(p.order.size() * 4));
  CUDA_TRY(cudaMemcpyAsync(e->d_order.p, p.order.data(), p.order.size() * 4, cudaMemcpyHostToDevice, e->st));
  CUDA_TRY(cudaMemsetAsync(e->d_counters.p, 0, 64, e->st));
  P.flanks = (const char *)e->d_flanks.p;
  P.gaps = (const GapDev *)e->d_gaps.p;
  P.outs = (GapOut *)e->d_outs.p;
  P.strs = (uint64_t *)e->d_chain_strs.p;
  P.walks = (ChainWalk *)e->d_chain_walks.p;
  P.fills = (char *)e->d_fill[p.pack_slot].p;
  P.fill_cap = p.fill_cap;
  P.fill_cursor = (unsigned long long *)((char *)e->d_counters.p + 16);
  CUDA_TRY(cudaEventRecord(e->ev[0], e->st));
  for (size_t at = 0; at < p.order.size(); at += chunk) {
    P.order = (const uint32_t *)e->d_order.p + at;
    P.n_gaps = (uint32_t)std::min(chunk, p.order.size() - at);
    CUDA_TRY(launch_chain(P, e->st));
    e->tm.launches += 2; e->tm.launches_chain += 2;
  }
  CUDA_TRY(cudaEventRecord(e->ev[1], e->st));
  CUDA_TRY(cudaMemcpyAsync(e->h_outs.p, e->d_outs.p, e->gaps.size() * sizeof(GapOut), cudaMemcpyDeviceToHost, e->st));
  unsigned long long fused = 0;
  CUDA_TRY(cudaMemcpyAsync(&fused, (char *)e->d_counters.p + 16, 8, cudaMemcpyDeviceToHost, e->st));
  CUDA_TRY(cudaStreamSynchronize(e->st));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
  e->tm.kernel_ms += ms; e->tm.kernel_ms_chain += ms;
  if (getenv("G2S_HOST_PROF"))
    fprintf(stderr, "g2s: chain tier: %zu gaps, strings %.1f MB, %.3f ms\n", p.order.size(), chunk * stride / 1e6, ms);
  e->tm.d2h_bytes += e->gaps.size() * sizeof(GapOut) + 8;
  p.pack_used = 0;
  p.fill_used = std::min(fused, p.fill_cap);
  return G2S_OK;
}

static int do_upload(g2s_engine *e, const g2s_gap *gaps, size_t n) {
  if (!e) return fail(G2S_E_ARG, "null engine");
  const bool multi = e->multi.d_tables != nullptr;
  if (!e->table.slots && !multi) return fail(G2S_E_STATE, "no graph imported (g2s_import_kmers / g2s_build_from_reads)");
  if (!gaps && n) return fail(G2S_E_ARG, "null gaps");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  if (multi) {   // per-gap graphs: g2s_gap.graph selects the gap's graph
    std::vector<uint32_t> graph_of(std::max<size_t>(n, 1), 0);
    for (size_t i = 0; i < n; i++) {
      if (gaps[i].graph < 0 || (uint64_t)gaps[i].graph >= e->multi.n_graphs) return fail(G2S_E_ARG, "g2s_gap.graph out of range");
      if (gaps[i].flags & G2S_GAP_PRESENT_BITS) return fail(G2S_E_ARG, "present_bits are not supported with per-gap graphs");
      graph_of[i] = (uint32_t)gaps[i].graph;
    }
    CUDA_TRY(e->d_graph_of.reserve(graph_of.size() * 4));
    CUDA_TRY(cudaMemcpyAsync(e->d_graph_of.p, graph_of.data(), graph_of.size() * 4, cudaMemcpyHostToDevice, e->st));
    CUDA_TRY(cudaStreamSynchronize(e->st));
  }
  const int k = e->cfg.k;
  e->tm = g2s_timing{};
  e->passes.clear();
  e->gaps.assign(gaps, gaps + n);
  e->gapdev.resize(n);
  e->outs.assign(n, GapOut{});
  e->out_pass.assign(n, -1);
  e->max_dmax = 0; e->max_rmf = 0;
  size_t fbytes = 0;
  for (size_t i = 0; i < n; i++) {
    const g2s_gap &g = gaps[i];
    if (!g.kmer_left || !g.kmer_right) return fail(G2S_E_ARG, "gap with null flank");
    if (g.gap_len < 0 || g.gap_err < 0 || g.left_max_fuz < 0 || g.right_max_fuz < 0 || g.left_max_fuz > e->cfg.max_fuz ||
        g.right_max_fuz > e->cfg.max_fuz)
      return fail(G2S_E_ARG, "gap parameters out of range (fuz must be <= g2s_config.max_fuz)");
    if ((long long)g.gap_len + g.gap_err > (1 << 24)) return fail(G2S_E_ARG, "gap_len + gap_err too large");
    fbytes += (size_t)2 * k + g.left_max_fuz + g.right_max_fuz + (g.kmer_right_starts ? (size_t)k + g.right_max_fuz : 0);
    e->max_dmax = std::max(e->max_dmax, dmax_of(g));
    e->max_rmf = std::max(e->max_rmf, std::max(g.right_max_fuz, g.left_max_fuz));
  }
  e->flank_store.resize(fbytes + 1);
  size_t off = 0;
  for (size_t i = 0; i < n; i++) {
    g2s_gap &g = e->gaps[i];
    GapDev &d = e->gapdev[i];
    d.flank_off = (uint32_t)off;
    memcpy(&e->flank_store[off], g.kmer_left, (size_t)k + g.left_max_fuz);
    g.kmer_left = nullptr;
    off += (size_t)k + g.left_max_fuz;
    d.rstart_off = (uint32_t)off;
    memcpy(&e->flank_store[off], g.kmer_right, (size_t)k + g.right_max_fuz);
    g.kmer_right = nullptr;
    off += (size_t)k + g.right_max_fuz;
    if (g.kmer_right_starts) {
      d.rstart_off = (uint32_t)off;
      memcpy(&e->flank_store[off], g.kmer_right_starts, (size_t)k + g.right_max_fuz);
      g.kmer_right_starts = nullptr;
      off += (size_t)k + g.right_max_fuz;
    }
    d.gap_len = g.gap_len; d.gap_err = g.gap_err; d.lmf = g.left_max_fuz; d.rmf = g.right_max_fuz;
    d.flags = (uint32_t)g.flags; d.id = (uint32_t)i;
  }
  if (fbytes >= (1ULL << 32)) return fail(G2S_E_ARG, "batch too large");
  CUDA_TRY(e->d_flanks.reserve(fbytes + 16));
  CUDA_TRY(e->d_gaps.reserve(std::max<size_t>(n, 1) * sizeof(GapDev)));
  CUDA_TRY(e->d_outs.reserve(std::max<size_t>(n, 1) * sizeof(GapOut)));
  CUDA_TRY(e->h_outs.reserve(std::max<size_t>(n, 1) * sizeof(GapOut)));
  CUDA_TRY(cudaEventRecord(e->ev[2], e->st));
  if (n) {
    CUDA_TRY(cudaMemcpyAsync(e->d_flanks.p, e->flank_store.data(), fbytes, cudaMemcpyHostToDevice, e->st));
    CUDA_TRY(cudaMemcpyAsync(e->d_gaps.p, e->gapdev.data(), n * sizeof(GapDev), cudaMemcpyHostToDevice, e->st));
  }
  CUDA_TRY(cudaEventRecord(e->ev[3], e->st));
  CUDA_TRY(cudaStreamSynchronize(e->st));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]));
  e->tm.h2d_ms = ms;
  e->tm.h2d_bytes = fbytes + n * sizeof(GapDev);
  // callers that keep their own graph hand in its contains() answers for the flank k-mers (INTEGRATION.md form 2): every
  // disagreement with the imported exact set is reported per gap, never silently resolved
  e->phantom.clear();
  bool any_bits = false;
  for (size_t i = 0; i < n && !any_bits; i++) any_bits = (gaps[i].flags & G2S_GAP_PRESENT_BITS) != 0;
  if (any_bits) {
    std::string ascii;
    std::vector<size_t> first(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
      first[i] = ascii.size() / k;
      if (!(gaps[i].flags & G2S_GAP_PRESENT_BITS)) continue;
      const GapDev &d = e->gapdev[i];
      const char *l = &e->flank_store[d.flank_off], *r = l + k + d.lmf, *rs = &e->flank_store[d.rstart_off];
      for (int q = 0; q <= d.lmf; q++) ascii.append(l + q, (size_t)k);
      for (int q = 0; q <= d.rmf; q++) ascii.append(r + q, (size_t)k);
      (void)rs;
    }
    first[n] = ascii.size() / k;
    std::vector<uint8_t> has(first[n] + 1, 0);
    CUDA_TRY(table_contains(e->table, ascii.data(), first[n], has.data(), e->st));
    e->phantom.assign(n, 0);
    for (size_t i = 0; i < n; i++) {
      if (!(gaps[i].flags & G2S_GAP_PRESENT_BITS)) continue;
      const GapDev &d = e->gapdev[i];
      size_t at = first[i];
      for (int q = 0; q <= d.lmf; q++, at++) if (((gaps[i].present_bits >> q) & 1) != (uint64_t)(has[at] != 0)) e->phantom[i] = 1;
      for (int q = 0; q <= d.rmf; q++, at++) if (((gaps[i].present_bits >> (32 + q)) & 1) != (uint64_t)(has[at] != 0)) e->phantom[i] = 1;
    }
  }
  e->uploaded = true; e->ran = false;
  return G2S_OK;
}

static int do_run(g2s_engine *e) {
  if (!e || !e->uploaded) return fail(G2S_E_STATE, "g2s_run_resident: nothing uploaded");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const size_t n = e->gaps.size();
  e->passes.clear();
  e->tm.kernel_ms = 0; e->tm.launches = 0; e->tm.gaps_big_tier = 0; e->tm.d2h_bytes = 0; e->tm.d2h_ms = 0; e->tm.host_ms = 0;
  e->tm.max_tier_run = 0; e->tm.gaps_rerun_scratch = 0; e->tm.gaps_rerun_out_full = 0; e->tm.gaps_mem_limit = 0;
  e->tm.table_lookups = 0; e->tm.node_ops = 0;
  e->tm.kernel_ms_thin = 0; e->tm.kernel_ms_wide = 0; e->tm.launches_thin = 0; e->tm.launches_wide = 0; e->tm.thin_shape = 0; e->tm.wide_cluster = 0;
  e->tm.cells_thin = 0; e->tm.cells_wide = 0; e->tm.kernel_ms_chain = 0; e->tm.launches_chain = 0; e->tm.gaps_chain = 0; e->tm.cells_chain = 0;
  e->out_pass.assign(n, -1);
  if (n == 0) { e->ran = true; return G2S_OK; }
  CUDA_TRY(cudaMemsetAsync(e->d_outs.p, 0, n * sizeof(GapOut), e->st));
  // Tiers: a gap that outgrows its scratch slice is run again in a 16x larger one.  How far it got tells how much
  // larger: a gap that overflowed before half of its levels skips a tier (on the repeat-rich workload nearly every gap
  // that fails the first tier also fails the second, which was a quarter of the device time).
  std::vector<int> want(n, 0);
  std::vector<uint8_t> done(n, 0);
  int max_tier = 0;
  { Pass probe; probe.order.push_back(0); for (int t = 1; t < 8; t++) { probe.tier = t; if (plan_pass(e, probe) != 0) break; max_tier = t; } }
  size_t pending = n;
  if (e->chain_tier) {   // gaps whose borders never exceed one state end here; the others go on to the tiers with tables
    e->passes.emplace_back();
    Pass &p = e->passes.back();
    p.tier = -1;
    p.pack_slot = 0;
    p.order.resize(n);
    for (size_t i = 0; i < n; i++) p.order[i] = (uint32_t)i;
    std::stable_sort(p.order.begin(), p.order.end(),
                     [&](uint32_t a, uint32_t b) { return dmax_of(e->gaps[a]) > dmax_of(e->gaps[b]); });
    int rc = run_chain_pass(e, p);
    if (rc != G2S_OK) return rc;
    const GapOut *ho = (const GapOut *)e->h_outs.p;
    for (size_t i = 0; i < n; i++) {
      if (!(ho[i].status & GS_DONE)) return fail(G2S_E_CUDA, "chain tier did not finish a gap");
      if (ho[i].status & GS_INTERNAL) return fail(G2S_E_CUDA, "device table inconsistent with its adjacency bytes");
      if (ho[i].status & GS_NOT_CHAIN) continue;
      e->outs[i] = ho[i]; e->out_pass[i] = 0; done[i] = 1; pending--;
    }
  }
  for (int tier = 0; pending > 0; tier++) {
    std::vector<uint32_t> todo;
    for (size_t i = 0; i < n; i++) if (!done[i] && want[i] <= tier) todo.push_back((uint32_t)i);
    if (todo.empty()) continue;
    if (tier >= 1 && tier < max_tier) {
      // A handful of gaps in this tier while most of the rest waits for the next one: they ride along there.  A pass
      // costs at least its slowest gap, so a small pass in front of a big one is pure latency (the repeat-rich workload:
      // 2 gaps in tier 1 took 1.5 ms before the 58 gaps of tier 2 could start); a larger slice changes no result.
      size_t next_cnt = 0;
      for (size_t i = 0; i < n; i++) if (!done[i] && want[i] == tier + 1) next_cnt++;
      if (next_cnt >= 4 * todo.size()) { for (uint32_t gi : todo) want[gi] = tier + 1; continue; }
    }
    e->passes.emplace_back();
    Pass &p = e->passes.back();
    p.tier = tier;
    p.pack_slot = (int)e->passes.size() - 1;
    p.order = todo;
    std::stable_sort(p.order.begin(), p.order.end(),
                     [&](uint32_t a, uint32_t b) { return dmax_of(e->gaps[a]) > dmax_of(e->gaps[b]); });
    if (tier > max_tier || plan_pass(e, p) != 0) {   // no larger tier fits the memory budget: the gaps are reported, not filled
      for (size_t i = 0; i < n; i++)
        if (!done[i]) { e->outs[i] = GapOut{}; e->outs[i].status = GS_OVERFLOW | GS_DONE; e->out_pass[i] = -2; e->tm.gaps_mem_limit++; }
      e->passes.pop_back();
      break;
    }
    int rc = run_pass(e, p);
    if (rc != G2S_OK) return rc;
    const GapOut *ho = (const GapOut *)e->h_outs.p;
    size_t failed = 0;
    for (uint32_t gi : todo) {
      if (!(ho[gi].status & GS_DONE)) return fail(G2S_E_CUDA, "device pass did not finish a gap");
      if (ho[gi].status & GS_INTERNAL) return fail(G2S_E_CUDA, "device table inconsistent with its adjacency bytes");
      if (ho[gi].status & (GS_OVERFLOW | GS_OUT_FULL)) {
        failed++;
        if (ho[gi].status & GS_OUT_FULL) e->tm.gaps_rerun_out_full++; else e->tm.gaps_rerun_scratch++;
        const g2s_gap &g = e->gaps[gi];
        const double levels = (double)dmax_of(g) + g.right_max_fuz + (g.gap_len + g.gap_err + 1) / 2;   // forward + right pass
        const bool early = !(ho[gi].status & GS_OUT_FULL) && (double)ho[gi].levels_r + ho[gi].levels_f < 0.5 * levels;
        want[gi] = std::min(tier + (early ? 2 : 1), std::max(max_tier, tier + 1));
      } else { e->outs[gi] = ho[gi]; e->out_pass[gi] = (int)e->passes.size() - 1; done[gi] = 1; pending--; }
    }
    if (tier > 0) e->tm.gaps_big_tier += (uint32_t)(todo.size() - failed);
    e->tm.max_tier_run = (uint32_t)tier;
  }
  for (size_t i = 0; i < n; i++) {
    const GapOut &o = e->outs[i];
    e->tm.table_lookups += o.table_lookups; e->tm.node_ops += o.node_ops;
    if (e->out_pass[i] < 0) continue;
    const int tier = e->passes[e->out_pass[i]].tier;
    (tier < 0 ? e->tm.cells_chain : tier == 0 ? e->tm.cells_thin : e->tm.cells_wide) += (uint64_t)o.cells_right + o.cells_left;
    if (tier < 0) e->tm.gaps_chain++;
  }
  e->ran = true;
  return G2S_OK;
}

// ------------------------------------------------------------------------------------------------ host stages
// D2H of the pack arenas + SCC analysis: everything of the batch that the host needs, detached from the engine
static int do_collect(g2s_engine *e, HostBatch &hb) {
  if (!e || !e->ran) return fail(G2S_E_STATE, "no device pass has run");
  CUDA_TRY(cudaSetDevice(e->cfg.device));
  const size_t n = e->gaps.size();
  hb.k = e->cfg.k;
  hb.gaps = e->gaps;
  hb.outs = e->outs;
  hb.tier.assign(n, -2);
  hb.phantom = e->phantom;
  hb.pack_of.assign(n, 0);
  hb.pack_base.assign(e->passes.size(), nullptr);
  hb.fill_base.assign(e->passes.size(), nullptr);
  auto bufs = std::make_shared<std::vector<PinBuf>>(2 * e->passes.size());   // [pass]: pack records, [passes + pass]: fills
  std::shared_ptr<PinPool> pool = e->pin;
  hb.on_free = [bufs, pool]() {   // hand the pinned staging back to the pool when the batch dies
    std::lock_guard<std::mutex> lk(pool->mu);
    for (auto &b : *bufs) if (b.p) pool->free.push_back(b);
    bufs->clear();
  };
  CUDA_TRY(cudaEventRecord(e->ev[2], e->st));
  for (size_t pi = 0; pi < e->passes.size(); pi++) {
    Pass &p = e->passes[pi];
    if (!p.pack_used) continue;
    const size_t need = p.pack_used * sizeof(PackRec);
    PinBuf &b = (*bufs)[pi];
    {
      std::lock_guard<std::mutex> lk(pool->mu);
      size_t best = pool->free.size();   // best fit: a small arena must not take the buffer a large one needs
      for (size_t j = 0; j < pool->free.size(); j++)
        if (pool->free[j].cap >= need && (best == pool->free.size() || pool->free[j].cap < pool->free[best].cap)) best = j;
      if (best < pool->free.size()) { b = pool->free[best]; pool->free.erase(pool->free.begin() + best); }
    }
    if (!b.p) {
      const auto ta = std::chrono::steady_clock::now();
      CUDA_TRY(b.reserve(need + need / 2));
      if (getenv("G2S_HOST_PROF"))
        fprintf(stderr, "g2s: pinned staging of %.1f MB allocated in %.3f ms\n", (need + need / 2) / 1e6,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ta).count());
    }
    CUDA_TRY(cudaMemcpyAsync(b.p, e->d_pack[p.pack_slot].p, need, cudaMemcpyDeviceToHost, e->st));
    hb.pack_base[pi] = (const PackRec *)b.p;
    hb.staging.emplace_back(b.p, need);
    e->tm.d2h_bytes += need;
  }
  for (size_t pi = 0; pi < e->passes.size(); pi++) {   // the fills the device traced itself
    Pass &p = e->passes[pi];
    if (!p.fill_used) continue;
    const size_t need = (size_t)p.fill_used;
    PinBuf &b = (*bufs)[e->passes.size() + pi];
    {
      std::lock_guard<std::mutex> lk(pool->mu);
      size_t best = pool->free.size();
      for (size_t j = 0; j < pool->free.size(); j++)
        if (pool->free[j].cap >= need && (best == pool->free.size() || pool->free[j].cap < pool->free[best].cap)) best = j;
      if (best < pool->free.size()) { b = pool->free[best]; pool->free.erase(pool->free.begin() + best); }
    }
    if (!b.p) CUDA_TRY(b.reserve(need + need / 2));
    CUDA_TRY(cudaMemcpyAsync(b.p, e->d_fill[p.pack_slot].p, need, cudaMemcpyDeviceToHost, e->st));
    hb.fill_base[pi] = (const char *)b.p;
    hb.staging.emplace_back(b.p, need);
    e->tm.d2h_bytes += need;
  }
  CUDA_TRY(cudaEventRecord(e->ev[3], e->st));
  CUDA_TRY(cudaStreamSynchronize(e->st));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]));
  e->tm.d2h_ms = ms;
  for (size_t i = 0; i < n; i++)
    if (e->out_pass[i] >= 0) { hb.tier[i] = e->passes[e->out_pass[i]].tier; hb.pack_of[i] = e->out_pass[i]; }
  auto t0 = std::chrono::steady_clock::now();
  hb.analyze(e->pool.get());
  e->tm.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (getenv("G2S_HOST_PROF")) fprintf(stderr, "g2s: subgraph analysis %.3f ms\n", e->tm.host_ms);
  return G2S_OK;
}

static int trace_batch(g2s_engine *e, HostBatch &hb, g2s_result *results, char *const *fill_out);
static int do_finish(g2s_engine *e, g2s_result *results, char *const *fill_out) {
  if (!e || !e->ran) return fail(G2S_E_STATE, "g2s_finish_resident: no device pass has run");
  if (!results && !e->gaps.empty()) return fail(G2S_E_ARG, "null results");
  HostBatch hb;
  int rc = do_collect(e, hb);
  if (rc != G2S_OK) return rc;
  return trace_batch(e, hb, results, fill_out);
}

// tracebacks of a whole batch in input order of the libc rand() stream, its only consumer (src/Gap2Seq.cpp:1440, 1513)
static int trace_batch(g2s_engine *e, HostBatch &hb, g2s_result *results, char *const *fill_out) {
  auto t0 = std::chrono::steady_clock::now();
  std::vector<TraceJob> jobs(hb.gaps.size());
  for (size_t i = 0; i < jobs.size(); i++) jobs[i] = TraceJob{&hb, i};
  std::vector<g2s_result> res;
  std::vector<size_t> off;
  if (e->pool && jobs.size() >= 4) e->ahead.begin();
  trace_many(jobs, res, e->trace_arena, off, e->pool.get(), e->ahead.ok() ? &e->ahead : nullptr);
  const char *arena = e->trace_arena.p;
  for (size_t i = 0; i < jobs.size(); i++) {
    results[i] = res[i];
    if (fill_out && fill_out[i]) memcpy(fill_out[i], &arena[off[i]], (size_t)res[i].fill_len + 1);
  }
  hb.flush_staging(e->pool.get());
  e->tm.host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return G2S_OK;
}


// ------------------------------------------------------------------------------------------------ one wave of gaps
// Longest-processing-time-first assignment of a wave's gaps to the GPUs: cost = levels of the DP
// (g + e + lmf + rmf) x the expected border width (1: no pilot pass here); the next most expensive gap goes to the
// GPU with the least load so far.  Each shard keeps input order.
static void lpt_shards(const g2s_gap *gaps, size_t n, int n_dev, std::vector<std::vector<uint32_t>> &shards) {
  shards.assign(n_dev, std::vector<uint32_t>());
  std::vector<uint32_t> order(n);
  for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return dmax_of(gaps[a]) > dmax_of(gaps[b]); });
  std::vector<double> load(n_dev, 0.0);
  for (uint32_t i : order) {
    int best = 0;
    for (int d = 1; d < n_dev; d++) if (load[d] < load[best]) best = d;
    shards[best].push_back(i);
    load[best] += (double)dmax_of(gaps[i]) + 64.0;
  }
  for (auto &sh : shards) std::sort(sh.begin(), sh.end());
}

static int do_upload(g2s_engine *e, const g2s_gap *gaps, size_t n);
static int do_run(g2s_engine *e);
static int do_collect(g2s_engine *e, HostBatch &hb);

// device pass + D2H + subgraph analysis of one wave of gaps -> a detached host batch in the wave's order (no rand()
// consumed).  With peers the wave is sharded over the GPUs, every GPU on its own host thread.
static int run_wave(g2s_engine *e, const g2s_gap *gaps, size_t n, HostBatch &hb, double *w_ms /* upload, run, collect */) {
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  if (e->peers.empty()) {
    auto t0 = now();
    int rc = do_upload(e, gaps, n);
    auto t1 = now();
    if (rc == G2S_OK) rc = do_run(e);
    auto t2 = now();
    if (rc == G2S_OK) rc = do_collect(e, hb);
    if (w_ms) { w_ms[0] += ms(t0, t1); w_ms[1] += ms(t1, t2); w_ms[2] += ms(t2, now()); }
    return rc;
  }
  const int nd = 1 + (int)e->peers.size();
  std::vector<g2s_engine *> eng(nd);
  eng[0] = e;
  for (int d = 1; d < nd; d++) eng[d] = e->peers[d - 1];
  for (int d = 1; d < nd; d++)
    if (!eng[d]->table.slots) return fail(G2S_E_STATE, "no graph on a peer GPU (import the graph through the multi-GPU engine)");
  std::vector<std::vector<uint32_t>> shards;
  lpt_shards(gaps, n, nd, shards);
  std::vector<std::shared_ptr<HostBatch>> parts(nd);
  std::vector<int> rcs(nd, G2S_OK);
  std::vector<std::string> errs(nd);
  std::vector<double> wall(3 * nd, 0.0);
  auto one_device = [&](int d) {
    std::vector<g2s_gap> mine(shards[d].size());
    for (size_t j = 0; j < mine.size(); j++) mine[j] = gaps[shards[d][j]];
    parts[d] = std::make_shared<HostBatch>();
    auto t0 = now();
    int rc = do_upload(eng[d], mine.data(), mine.size());
    auto t1 = now();
    if (rc == G2S_OK) rc = do_run(eng[d]);
    auto t2 = now();
    if (rc == G2S_OK) rc = do_collect(eng[d], *parts[d]);
    wall[3 * d] = ms(t0, t1); wall[3 * d + 1] = ms(t1, t2); wall[3 * d + 2] = ms(t2, now());
    rcs[d] = rc;
    if (rc != G2S_OK) errs[d] = g_err;
  };
  {
    std::vector<std::thread> th;
    for (int d = 1; d < nd; d++) th.emplace_back(one_device, d);
    one_device(0);
    for (auto &t : th) t.join();
  }
  cudaSetDevice(e->cfg.device);
  for (int d = 0; d < nd; d++) if (rcs[d] != G2S_OK) return fail(rcs[d], "GPU " + std::to_string(eng[d]->cfg.device) + ": " + errs[d]);
  // merge in the wave's order; the parts stay alive (and their pinned staging busy) as long as the merged batch
  hb.k = e->cfg.k;
  hb.gaps.resize(n); hb.outs.resize(n); hb.tier.assign(n, -2); hb.pack_of.assign(n, 0); hb.gen.resize(n);
  hb.pack_base.clear(); hb.fill_base.clear();
  for (int d = 0; d < nd; d++) {
    HostBatch &pt = *parts[d];
    const int base = (int)hb.pack_base.size();
    hb.pack_base.insert(hb.pack_base.end(), pt.pack_base.begin(), pt.pack_base.end());
    hb.fill_base.insert(hb.fill_base.end(), pt.fill_base.begin(), pt.fill_base.end());
    hb.staging.insert(hb.staging.end(), pt.staging.begin(), pt.staging.end());
    pt.staging.clear();
    for (size_t j = 0; j < shards[d].size(); j++) {
      const uint32_t i = shards[d][j];
      hb.gaps[i] = pt.gaps[j]; hb.outs[i] = pt.outs[j]; hb.tier[i] = pt.tier[j]; hb.pack_of[i] = base + pt.pack_of[j];
      if (j < pt.phantom.size() && pt.phantom[j]) { if (hb.phantom.empty()) hb.phantom.assign(n, 0); hb.phantom[i] = 1; }
      if (j < pt.gen.size()) hb.gen[i] = std::move(pt.gen[j]);
    }
  }
  hb.parts = parts;
  // timing of the wave: the GPUs ran side by side
  g2s_timing t{};
  for (int d = 0; d < nd; d++) {
    const g2s_timing &x = eng[d]->tm;
    t.h2d_ms = std::max(t.h2d_ms, x.h2d_ms); t.kernel_ms = std::max(t.kernel_ms, x.kernel_ms); t.d2h_ms = std::max(t.d2h_ms, x.d2h_ms);
    t.host_ms = std::max(t.host_ms, x.host_ms);
    t.h2d_bytes += x.h2d_bytes; t.d2h_bytes += x.d2h_bytes; t.launches += x.launches; t.gaps_big_tier += x.gaps_big_tier;
    t.max_tier_run = std::max(t.max_tier_run, x.max_tier_run); t.gaps_rerun_scratch += x.gaps_rerun_scratch;
    t.gaps_rerun_out_full += x.gaps_rerun_out_full; t.gaps_mem_limit += x.gaps_mem_limit;
    t.table_lookups += x.table_lookups; t.node_ops += x.node_ops;
    t.kernel_ms_thin = std::max(t.kernel_ms_thin, x.kernel_ms_thin); t.kernel_ms_wide = std::max(t.kernel_ms_wide, x.kernel_ms_wide);
    t.launches_thin += x.launches_thin; t.launches_wide += x.launches_wide; t.thin_shape = std::max(t.thin_shape, x.thin_shape); t.wide_cluster = std::max(t.wide_cluster, x.wide_cluster);
    t.cells_thin += x.cells_thin; t.cells_wide += x.cells_wide; t.cells_chain += x.cells_chain; t.gaps_chain += x.gaps_chain;
    t.kernel_ms_chain = std::max(t.kernel_ms_chain, x.kernel_ms_chain); t.launches_chain += x.launches_chain;
    if (getenv("G2S_HOST_PROF"))
      fprintf(stderr, "g2s: GPU %d: %zu gaps, upload %.3f ms, run %.3f ms (kernels %.3f), collect %.3f ms\n", eng[d]->cfg.device,
              shards[d].size(), wall[3 * d], wall[3 * d + 1], x.kernel_ms, wall[3 * d + 2]);
  }
  if (w_ms)
    for (int q = 0; q < 3; q++) { double m = 0; for (int d = 0; d < nd; d++) m = std::max(m, wall[3 * d + q]); w_ms[q] += m; }
  e->tm = t;
  return G2S_OK;
}

// ------------------------------------------------------------------------------------------------ C ABI: fill
extern "C" int g2s_upload_gaps(g2s_engine *e, const g2s_gap *gaps, size_t n) { return do_upload(e, gaps, n); }
extern "C" int g2s_run_resident(g2s_engine *e) { return do_run(e); }
extern "C" int g2s_finish_resident(g2s_engine *e, g2s_result *results, char *const *fill_out) { return do_finish(e, results, fill_out); }

extern "C" int g2s_fill_batch(g2s_engine *e, const g2s_gap *gaps, size_t n, g2s_result *results, char *const *fill_out) {
  if (!e) return fail(G2S_E_ARG, "null engine");
  if (e->peers.empty()) {
    int rc = do_upload(e, gaps, n);
    if (rc == G2S_OK) rc = do_run(e);
    if (rc == G2S_OK) rc = do_finish(e, results, fill_out);
    return rc;
  }
  if (!results && n) return fail(G2S_E_ARG, "null results");
  HostBatch hb;
  int rc = run_wave(e, gaps, n, hb, nullptr);
  if (rc != G2S_OK) return rc;
  return trace_batch(e, hb, results, fill_out);
}

struct g2s_batch { HostBatch hb; };

extern "C" int g2s_submit(g2s_engine *e, const g2s_gap *gaps, size_t n, g2s_batch **out) {
  if (!out || !e) return fail(G2S_E_ARG, "g2s_submit: null argument");
  g2s_batch *b = new g2s_batch();
  int rc = run_wave(e, gaps, n, b->hb, nullptr);
  if (rc != G2S_OK) { delete b; return rc; }
  *out = b;
  return G2S_OK;
}

extern "C" int g2s_batch_peek(const g2s_batch *b, size_t i, int32_t *count, int32_t *right_fuz) {
  if (!b || i >= b->hb.gaps.size()) return fail(G2S_E_ARG, "g2s_batch_peek: bad index");
  if (count) *count = b->hb.peek_count(i);
  if (right_fuz) *right_fuz = b->hb.peek_right_fuz(i);
  return G2S_OK;
}

extern "C" int g2s_batch_trace(g2s_batch *b, size_t i, g2s_result *result, char *fill_out) {
  if (!b || !result || i >= b->hb.gaps.size()) return fail(G2S_E_ARG, "g2s_batch_trace: bad argument");
  b->hb.trace(i, *result, fill_out);
  return G2S_OK;
}

extern "C" void g2s_batch_free(g2s_batch *b) { delete b; }
extern "C" void g2s_free(void *p) { free(p); }

static int fill_records_impl(g2s_engine *e, const char *const *seqs, const char *const *comments, size_t n_records,
                                const g2s_run_options *opt, char **filled_out, char **log_out, int32_t *gapcount,
                                int32_t *filledcount, g2s_result **gap_results, size_t *n_gap_results) {
  if (!e || !opt || (n_records && (!seqs || !comments || !filled_out))) return fail(G2S_E_ARG, "g2s_fill_records: null argument");
  if (opt->k != e->cfg.k) return fail(G2S_E_ARG, "g2s_fill_records: k differs from the engine's k");
  if (opt->fuz < 0 || opt->fuz > e->cfg.max_fuz) return fail(G2S_E_ARG, "g2s_fill_records: fuz exceeds g2s_config.max_fuz");
  for (size_t i = 0; i < n_records; i++) filled_out[i] = nullptr;
  std::vector<StrRef> s(n_records), c(n_records);   // the caller's strings, not copied
  auto ref_one = [&](size_t i) { s[i] = StrRef(seqs[i]); c[i] = StrRef(comments[i]); };
  if (e->pool) e->pool->parallel_for(n_records, 64, ref_one);
  else for (size_t i = 0; i < n_records; i++) ref_one(i);
  RunOptions ro;
  ro.k = opt->k; ro.d_err = opt->dist_error; ro.max_fuz = opt->fuz;
  ro.skip_confident = opt->skip_confident != 0; ro.unique_paths = opt->unique_paths != 0; ro.all_paths = opt->all_paths != 0;
  ScaffoldRun run(s, c, ro);
  run.pool = e->pool.get();
  run.aheadp = &e->ahead;
  run.arenap = &e->trace_arena;
  run.filled_c = filled_out;   // the records are written straight into the malloc'ed strings the caller receives
  g2s_timing total{};
  int err = G2S_OK;
  double w_ms[3] = {0, 0, 0};
  double &w_upload = w_ms[0], &w_run = w_ms[1], &w_collect = w_ms[2];
  SubmitFn submit = [&](const std::vector<g2s_gap> &gaps, HostBatch &hb) -> int {
    int rc = run_wave(e, gaps.data(), gaps.size(), hb, w_ms);
    err = rc;
    total.h2d_ms += e->tm.h2d_ms; total.kernel_ms += e->tm.kernel_ms; total.d2h_ms += e->tm.d2h_ms; total.host_ms += e->tm.host_ms;
    total.h2d_bytes += e->tm.h2d_bytes; total.d2h_bytes += e->tm.d2h_bytes; total.launches += e->tm.launches;
    total.gaps_big_tier += e->tm.gaps_big_tier;
    total.max_tier_run = std::max(total.max_tier_run, e->tm.max_tier_run);
    total.gaps_rerun_scratch += e->tm.gaps_rerun_scratch; total.gaps_rerun_out_full += e->tm.gaps_rerun_out_full;
    total.gaps_mem_limit += e->tm.gaps_mem_limit;
    total.table_lookups += e->tm.table_lookups; total.node_ops += e->tm.node_ops;
    total.kernel_ms_thin += e->tm.kernel_ms_thin; total.kernel_ms_wide += e->tm.kernel_ms_wide;
    total.launches_thin += e->tm.launches_thin; total.launches_wide += e->tm.launches_wide; total.thin_shape = e->tm.thin_shape; total.wide_cluster = std::max(total.wide_cluster, e->tm.wide_cluster);
    total.cells_thin += e->tm.cells_thin; total.cells_wide += e->tm.cells_wide; total.cells_chain += e->tm.cells_chain;
    total.gaps_chain += e->tm.gaps_chain; total.kernel_ms_chain += e->tm.kernel_ms_chain; total.launches_chain += e->tm.launches_chain;
    return rc;
  };
  std::vector<std::string> filled;
  std::string log;
  int gc = 0, fc = 0;
  std::vector<g2s_result> details;
  auto t0 = std::chrono::steady_clock::now();
  int rc = run.run(submit, filled, log, gc, fc, gap_results ? &details : nullptr);
  if (rc != 0) return err != G2S_OK ? err : fail(G2S_E_STATE, "g2s_fill_records: scanner inconsistency " + std::to_string(rc));
  const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  total.host_ms = std::max(0.0, wall - total.h2d_ms - total.kernel_ms - total.d2h_ms);
  e->tm = total;
  if (getenv("G2S_HOST_PROF"))   // profiling aid: where the wall time of this call went
    fprintf(stderr, "g2s_fill_records: wall %.3f ms = scan+flanks %.3f | upload %.3f | run %.3f (kernel %.3f) | collect %.3f "
            "(d2h %.3f) | trace %.3f (rand wait %.3f, flush %.3f) | splice+log %.3f\n", wall, run.phase1_ms - w_upload - w_run - w_collect, w_upload, w_run,
            total.kernel_ms, w_collect, total.d2h_ms, run.trace_ms, run.wait_ms, run.flush_ms, run.phase2_ms - run.trace_ms);
  for (size_t i = 0; i < n_records; i++)
    if (!filled_out[i]) return fail(G2S_E_NOMEM, "g2s_fill_records: out of host memory");
  if (log_out) *log_out = strdup(log.c_str());
  if (gapcount) *gapcount = gc;
  if (filledcount) *filledcount = fc;
  if (gap_results && n_gap_results) {
    *n_gap_results = details.size();
    *gap_results = (g2s_result *)malloc(std::max<size_t>(details.size(), 1) * sizeof(g2s_result));
    memcpy(*gap_results, details.data(), details.size() * sizeof(g2s_result));
  }
  return G2S_OK;
}

extern "C" int g2s_fill_records(g2s_engine *e, const char *const *seqs, const char *const *comments, size_t n_records,
                                const g2s_run_options *opt, char **filled_out, char **log_out, int32_t *gapcount,
                                int32_t *filledcount, g2s_result **gap_results, size_t *n_gap_results) {
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = fill_records_impl(e, seqs, comments, n_records, opt, filled_out, log_out, gapcount, filledcount, gap_results,
                                   n_gap_results);
  if (rc != G2S_OK && filled_out)   // nothing half-written is handed out
    for (size_t i = 0; i < n_records; i++) { free(filled_out[i]); filled_out[i] = nullptr; }
  if (getenv("G2S_HOST_PROF"))
    fprintf(stderr, "g2s_fill_records: %.3f ms from entry to return (teardown of the run's host state included)\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  return rc;
}

extern "C" int g2s_fill_gap_ex(g2s_engine *e, const char *kmer_left, const char *kmer_right, int gap_len, int k, int gap_err,
                               int left_max_fuz, int right_max_fuz, int *left_fuz, int *right_fuz, long long max_mem, char *fill,
                               int skip_confident, int all_paths, g2s_subgraph_stats *substats, uint64_t present_bits,
                               int have_present_bits, int32_t *status) {
  (void)max_mem;   // the device tiers are bounded by g2s_config.max_mem_bytes
  if (!e) return fail(G2S_E_ARG, "null engine");
  if (k != e->cfg.k) return fail(G2S_E_ARG, "g2s_fill_gap: k differs from the engine's k");
  if (!left_fuz || !right_fuz || !fill) return fail(G2S_E_ARG, "g2s_fill_gap: null output");
  g2s_gap g{};
  g.kmer_left = kmer_left; g.kmer_right = kmer_right;
  const size_t rlen = strlen(kmer_right);
  if (strlen(kmer_left) < (size_t)k + left_max_fuz || rlen < (size_t)k + right_max_fuz)
    return fail(G2S_E_ARG, "g2s_fill_gap: flank shorter than k + max_fuz");
  if (rlen > (size_t)k + right_max_fuz) g.kmer_right_starts = kmer_right + (rlen - k - right_max_fuz);
  g.gap_len = gap_len; g.gap_err = gap_err; g.left_max_fuz = left_max_fuz; g.right_max_fuz = right_max_fuz;
  g.flags = (skip_confident ? G2S_GAP_SKIP_CONFIDENT : 0) | (all_paths ? G2S_GAP_ALL_PATHS : 0) |
            (have_present_bits ? G2S_GAP_PRESENT_BITS : 0);
  g.present_bits = present_bits;
  g2s_result r;
  std::vector<char> tmp((size_t)gap_len + k + gap_err + left_max_fuz + right_max_fuz + 3);
  char *fo = tmp.data();
  int rc = g2s_fill_batch(e, &g, 1, &r, &fo);
  if (rc != G2S_OK) return rc;
  if (r.fill_len > 0 || r.count > 0) {
    *left_fuz = r.left_fuz; *right_fuz = r.right_fuz;
    memcpy(fill + (left_max_fuz - r.left_fuz), tmp.data(), (size_t)r.fill_len + 1);   // same offsets as the reference
  } else if (r.forward_count > 0) {
    *right_fuz = r.right_fuz;
  }
  if (substats && r.substats_valid) *substats = r.substats;
  if (status) *status = r.status;
  return r.count;
}

extern "C" int g2s_fill_gap(g2s_engine *e, const char *kmer_left, const char *kmer_right, int gap_len, int k, int gap_err,
                            int left_max_fuz, int right_max_fuz, int *left_fuz, int *right_fuz, long long max_mem, char *fill,
                            int skip_confident, int all_paths, g2s_subgraph_stats *substats) {
  return g2s_fill_gap_ex(e, kmer_left, kmer_right, gap_len, k, gap_err, left_max_fuz, right_max_fuz, left_fuz, right_fuz, max_mem,
                         fill, skip_confident, all_paths, substats, 0, 0, nullptr);
}
