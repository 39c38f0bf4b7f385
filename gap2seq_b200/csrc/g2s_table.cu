// g2s_table.cu - the device-resident de Bruijn graph: an open-addressing table of canonical k-mers with one
// adjacency byte per node.  Replaces the query side of GATB's Graph (Bloom filter + cFP set + MPHF,
// gatb-core/src/gatb/debruijn/impl/Graph.hpp:1249-1272; neighbours by shift/insert + 4 membership tests,
// Graph.cpp:1284-1345) and, for stand-alone use, its construction (DSK solid k-mer counting invoked from
// src/Gap2Seq.cpp:213-214) by sort + run-length counting on the device.
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "g2s_common.cuh"
#include "g2s_device.h"

namespace g2s {

// ---------------------------------------------------------------------------------------------- table build
__global__ void k_clear_slots(Slot *slots, uint64_t cap) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < cap; i += stride) { slots[i].key = KEY_EMPTY; slots[i].val = 0; slots[i].h[0] = 0; slots[i].h[1] = 0; }
}

// insert keys[i] when its abundance passes; counts == nullptr means "insert all"
__global__ void k_insert(Slot *slots, uint64_t cap_mask, const uint64_t *keys, const uint32_t *counts, uint32_t solid,
                         uint64_t n, unsigned long long *n_inserted) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long mine = 0;
  for (; i < n; i += stride) {
    const uint64_t key = keys[i];
    if (key == KEY_EMPTY) continue;
    if (counts && counts[i] < solid) continue;
    uint64_t h = mix64(key) & cap_mask;
    for (;;) {
      unsigned long long old = atomicCAS((unsigned long long *)&slots[h].key, (unsigned long long)KEY_EMPTY,
                                         (unsigned long long)key);
      if (old == KEY_EMPTY) { mine++; break; }
      if (old == key) break;
      h = (h + 1) & cap_mask;
    }
  }
  // one atomic per warp
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_inserted, mine);
}

// adjacency byte of every node: 8 probes (4 successors, 4 predecessors of the canonical orientation),
// the same byte Graph::precomputeAdjacency computes (Graph.cpp:3486-3530)
__global__ void k_adjacency(Slot *slots, Table t) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t cap = t.cap_mask + 1;
  for (; i < cap; i += stride) {
    const uint64_t key = slots[i].key;
    if (key == KEY_EMPTY) continue;
    uint32_t adj = 0, dummy;
#pragma unroll
    for (uint32_t nt = 0; nt < 4; nt++) {
      if (table_find(t, canonical(succ_kmer(key, nt, t), t.k).key, &dummy) >= 0) adj |= 1u << nt;
      if (table_find(t, canonical(pred_kmer(key, nt, t), t.k).key, &dummy) >= 0) adj |= 1u << (4 + nt);
    }
    slots[i].val = adj;
  }
}

// skip hints of every node (g2s_common.cuh hint_of): up to 40 dependent probes per orientation, run once at import
__global__ void k_hints(const Slot *slots, Table t, Slot *hints) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t cap = t.cap_mask + 1;
  for (; i < cap; i += stride) {
    const uint64_t key = slots[i].key;
    Slot o{key, 0, {0, 0}};
    if (key != KEY_EMPTY) compute_hints(t, key, (uint32_t)(slots[i].val & 0xFF), &o);
    hints[i] = o;
  }
}

// second step so that k_hints reads a table nobody writes (its probes use the read-only path)
__global__ void k_apply_hints(Slot *slots, uint64_t cap, const Slot *hints) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < cap; i += stride)
    if (slots[i].key != KEY_EMPTY) { slots[i].val = hints[i].val; slots[i].h[0] = hints[i].h[0]; slots[i].h[1] = hints[i].h[1]; }
}

__global__ void k_count_solid(const uint32_t *counts, const int *n_runs, const uint64_t *keys, uint32_t solid,
                              unsigned long long *n_solid) {
  const uint64_t n = (uint64_t)*n_runs;
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long mine = 0;
  for (; i < n; i += stride) mine += (counts[i] >= solid && keys[i] != KEY_EMPTY) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_solid, mine);
}

// one thread per (read, position): canonical k-mer or KEY_EMPTY when the window holds an invalid character
// ((c >> 3) & 1, tools/misc/api/Data.hpp:176; such k-mers are not counted, kmer/impl/Sequence2SuperKmer.hpp:93)
__global__ void k_extract(const char *bases, const uint64_t *offsets, const uint64_t *kmer_offsets, uint64_t n_reads,
                          int k, uint64_t *out, uint64_t total) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    // binary search the read that owns output index i
    uint64_t lo = 0, hi = n_reads;
    while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (kmer_offsets[mid] <= i) lo = mid; else hi = mid; }
    const char *s = bases + offsets[lo] + (i - kmer_offsets[lo]);
    uint64_t v = 0; uint32_t bad = 0;
    for (int j = 0; j < k; j++) {
      const unsigned char c = (unsigned char)s[j];
      bad |= (c >> 3) & 1;
      v = (v << 2) | ((c >> 1) & 3);
    }
    out[i] = bad ? KEY_EMPTY : canonical(v, k).key;
  }
}

// Partitioned counting (inputs with 2^31 or more k-mer occurrences, or more than the sort buffers should hold): the
// occurrences are split by a hash of the canonical k-mer, so every partition is counted on its own with all occurrences
// of its k-mers present.  part = (mix64(key) >> 40) & pmask.
__device__ __forceinline__ uint64_t kmer_at(const char *bases, const uint64_t *offsets, const uint64_t *kmer_offsets,
                                            uint64_t n_reads, int k, uint64_t i) {
  uint64_t lo = 0, hi = n_reads;
  while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (kmer_offsets[mid] <= i) lo = mid; else hi = mid; }
  const char *s = bases + offsets[lo] + (i - kmer_offsets[lo]);
  uint64_t v = 0; uint32_t bad = 0;
  for (int j = 0; j < k; j++) {
    const unsigned char c = (unsigned char)s[j];
    bad |= (c >> 3) & 1;
    v = (v << 2) | ((c >> 1) & 3);
  }
  return bad ? KEY_EMPTY : canonical(v, k).key;
}

__global__ void k_part_histogram(const char *bases, const uint64_t *offsets, const uint64_t *kmer_offsets, uint64_t n_reads,
                                 int k, uint64_t total, uint32_t pmask, unsigned long long *hist) {
  __shared__ unsigned long long sh[256];
  for (uint32_t q = threadIdx.x; q <= pmask; q += blockDim.x) sh[q] = 0;
  __syncthreads();
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const uint64_t key = kmer_at(bases, offsets, kmer_offsets, n_reads, k, i);
    if (key != KEY_EMPTY) atomicAdd(&sh[(uint32_t)(mix64(key) >> 40) & pmask], 1ULL);
  }
  __syncthreads();
  for (uint32_t q = threadIdx.x; q <= pmask; q += blockDim.x) if (sh[q]) atomicAdd(&hist[q], sh[q]);
}

__global__ void k_extract_part(const char *bases, const uint64_t *offsets, const uint64_t *kmer_offsets, uint64_t n_reads,
                               int k, uint64_t total, uint32_t pmask, uint32_t part, uint64_t *out, unsigned long long *cursor) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const uint64_t key = kmer_at(bases, offsets, kmer_offsets, n_reads, k, i);
    if (key != KEY_EMPTY && ((uint32_t)(mix64(key) >> 40) & pmask) == part) out[atomicAdd(cursor, 1ULL)] = key;
  }
}

// unique k-mers of one partition with count >= solid, appended to the list of solid k-mers
__global__ void k_append_solid(const uint64_t *keys, const uint32_t *counts, const int *n_runs, uint32_t solid,
                               uint64_t *solid_keys, unsigned long long *cursor) {
  const uint64_t n = (uint64_t)*n_runs;
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride)
    if (counts[i] >= solid && keys[i] != KEY_EMPTY) solid_keys[atomicAdd(cursor, 1ULL)] = keys[i];
}

__global__ void k_compact(const Slot *slots, uint64_t cap, uint64_t *keys_out, unsigned long long *cursor) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < cap; i += stride) {
    const uint64_t key = slots[i].key;
    if (key != KEY_EMPTY) keys_out[atomicAdd(cursor, 1ULL)] = key;
  }
}

__global__ void k_lookup_adj(Table t, const uint64_t *keys, uint64_t n, uint8_t *adj_out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t adj = 0;
  adj_out[i] = table_find(t, keys[i], &adj) >= 0 ? (uint8_t)adj : 0;
}

__global__ void k_contains(Table t, const char *ascii, uint64_t n, uint8_t *out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t adj;
  out[i] = table_find(t, canonical(kmer_from_ascii(ascii + i * t.k, t.k), t.k).key, &adj) >= 0;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

static int grid_for(uint64_t n, int block, int sms) {
  uint64_t b = (n + block - 1) / block;
  uint64_t cap = (uint64_t)sms * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t table_alloc(DeviceTable *dt, int k, uint64_t n_keys, int sms, cudaStream_t st) {
  uint64_t cap = 1024;
  while (cap < 2 * n_keys + 2) cap <<= 1;
  if (dt->slots) { cudaFree(dt->slots); dt->slots = nullptr; }
  CK(cudaMalloc(&dt->slots, cap * sizeof(Slot)));
  dt->cap = cap;
  dt->k = k;
  dt->n = 0;
  k_clear_slots<<<grid_for(cap, 256, sms), 256, 0, st>>>(dt->slots, cap);
  return cudaGetLastError();
}

Table table_view(const DeviceTable &dt) {
  Table t;
  t.slots = dt.slots;
  t.cap_mask = dt.cap - 1;
  t.k = dt.k;
  t.kmask = dt.k >= 32 ? ~0ULL : ((1ULL << (2 * dt.k)) - 1);
  return t;
}

// keys_dev: device array of canonical k-mers (counts_dev optional)
cudaError_t table_fill(DeviceTable *dt, const uint64_t *keys_dev, const uint32_t *counts_dev, uint32_t solid, uint64_t n,
                       int sms, cudaStream_t st) {
  unsigned long long *d_n;
  CK(cudaMalloc(&d_n, sizeof(unsigned long long)));
  CK(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), st));
  k_insert<<<grid_for(n, 256, sms), 256, 0, st>>>(dt->slots, dt->cap - 1, keys_dev, counts_dev, solid, n, d_n);
  CK(cudaGetLastError());
  k_adjacency<<<grid_for(dt->cap, 256, sms), 256, 0, st>>>(dt->slots, table_view(*dt));
  CK(cudaGetLastError());
  Slot *d_hints = nullptr;
  CK(cudaMalloc(&d_hints, dt->cap * sizeof(Slot)));
  k_hints<<<grid_for(dt->cap, 256, sms), 256, 0, st>>>(dt->slots, table_view(*dt), d_hints);
  CK(cudaGetLastError());
  k_apply_hints<<<grid_for(dt->cap, 256, sms), 256, 0, st>>>(dt->slots, dt->cap, d_hints);
  CK(cudaGetLastError());
  unsigned long long h_n = 0;
  CK(cudaMemcpyAsync(&h_n, d_n, sizeof(h_n), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(d_n);
  cudaFree(d_hints);
  dt->n = h_n;
  return cudaSuccess;
}

// counting core on device-resident reads: d_bases (concatenated bases), d_off / d_koff (n_reads + 1 exclusive offsets of
// the reads' bases / k-mer occurrences), total = k-mer occurrences
static cudaError_t table_from_device_reads(DeviceTable *dt, int k, const char *d_bases, const uint64_t *d_off, const uint64_t *d_koff,
                                           uint64_t n_reads, uint64_t total, int solid, int sms, cudaStream_t st) {
  uint64_t *d_keys = nullptr, *d_sorted = nullptr;
  uint64_t *d_unique = nullptr; uint32_t *d_counts = nullptr; int *d_nruns = nullptr; void *d_tmp = nullptr;
  unsigned long long *d_nsolid = nullptr;
  cudaError_t err = cudaSuccess;
#define CKG(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)
  {
    if (total == 0) { CKG(table_alloc(dt, k, 0, sms, st)); CKG(cudaStreamSynchronize(st)); goto done; }
    // One sort + run-length pass holds at most `limit` occurrences (cub's run-length API counts in int; the sort needs two
    // 8-byte buffers per occurrence).  Larger inputs are counted in hash partitions of the canonical k-mer.
    uint64_t limit = 1ULL << 30;
    if (const char *lv = getenv("G2S_COUNT_PART_LIMIT")) { const uint64_t v = strtoull(lv, nullptr, 10); if (v >= 1024) limit = v; }
    if (total > limit) {
      uint32_t pbits = 1;
      while (pbits < 8 && (total >> pbits) > limit - limit / 4) pbits++;   // a quarter of headroom for uneven partitions
      const uint32_t nparts = 1u << pbits, pmask = nparts - 1;
      unsigned long long *d_hist = nullptr, *d_cur = nullptr;
      uint64_t *d_solid = nullptr;
      unsigned long long hist[256];
      CKG(cudaMalloc(&d_hist, nparts * 8));
      CKG(cudaMemsetAsync(d_hist, 0, nparts * 8, st));
      k_part_histogram<<<grid_for(total, 256, sms), 256, 0, st>>>(d_bases, d_off, d_koff, n_reads, k, total, pmask, d_hist);
      err = cudaGetLastError();
      if (err == cudaSuccess) err = cudaMemcpyAsync(hist, d_hist, nparts * 8, cudaMemcpyDeviceToHost, st);
      if (err == cudaSuccess) err = cudaStreamSynchronize(st);
      cudaFree(d_hist);
      if (err != cudaSuccess) goto done;
      uint64_t biggest = 0, valid = 0;
      for (uint32_t q = 0; q < nparts; q++) { biggest = std::max<uint64_t>(biggest, hist[q]); valid += hist[q]; }
      if (biggest >= (1ULL << 31)) { err = cudaErrorInvalidValue; goto done; }   // 256 partitions of 2^31: beyond this build
      const uint64_t solid_cap = valid / (uint64_t)std::max(solid, 1) + 1;          // a solid k-mer has >= solid occurrences
      CKG(cudaMalloc(&d_keys, std::max<uint64_t>(biggest, 1) * 8));
      CKG(cudaMalloc(&d_sorted, std::max<uint64_t>(biggest, 1) * 8));
      CKG(cudaMalloc(&d_counts, std::max<uint64_t>(biggest, 1) * 4));
      CKG(cudaMalloc(&d_nruns, sizeof(int)));
      CKG(cudaMalloc(&d_solid, solid_cap * 8));
      CKG(cudaMalloc(&d_cur, 16));
      CKG(cudaMemsetAsync(d_cur, 0, 16, st));
      for (uint32_t q = 0; q < nparts && err == cudaSuccess; q++) {
        if (hist[q] == 0) continue;
        err = cudaMemsetAsync(d_cur, 0, 8, st);
        if (err != cudaSuccess) break;
        k_extract_part<<<grid_for(total, 256, sms), 256, 0, st>>>(d_bases, d_off, d_koff, n_reads, k, total, pmask, q, d_keys, d_cur);
        size_t tmp_bytes = 0;
        err = cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_sorted, (int)hist[q], 0, 64, st);
        if (err != cudaSuccess) break;
        err = cudaMalloc(&d_tmp, tmp_bytes);
        if (err != cudaSuccess) break;
        err = cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_sorted, (int)hist[q], 0, 64, st);
        cudaFree(d_tmp); d_tmp = nullptr;
        if (err != cudaSuccess) break;
        tmp_bytes = 0;
        err = cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, d_sorted, d_keys, d_counts, d_nruns, (int)hist[q], st);
        if (err != cudaSuccess) break;
        err = cudaMalloc(&d_tmp, tmp_bytes);
        if (err != cudaSuccess) break;
        err = cub::DeviceRunLengthEncode::Encode(d_tmp, tmp_bytes, d_sorted, d_keys, d_counts, d_nruns, (int)hist[q], st);
        cudaFree(d_tmp); d_tmp = nullptr;
        if (err != cudaSuccess) break;
        k_append_solid<<<grid_for(hist[q], 256, sms), 256, 0, st>>>(d_keys, d_counts, d_nruns, (uint32_t)solid, d_solid, d_cur + 1);
        err = cudaGetLastError();
      }
      unsigned long long n_solid = 0;
      if (err == cudaSuccess) err = cudaMemcpyAsync(&n_solid, d_cur + 1, 8, cudaMemcpyDeviceToHost, st);
      if (err == cudaSuccess) err = cudaStreamSynchronize(st);
      if (err == cudaSuccess) err = table_alloc(dt, k, n_solid, sms, st);
      if (err == cudaSuccess) err = table_fill(dt, d_solid, nullptr, 0, n_solid, sms, st);
      cudaFree(d_solid); cudaFree(d_cur);
      goto done;
    }
    CKG(cudaMalloc(&d_keys, total * 8));
    CKG(cudaMalloc(&d_sorted, total * 8));
    k_extract<<<grid_for(total, 256, sms), 256, 0, st>>>(d_bases, d_off, d_koff, n_reads, k, d_keys, total);
    CKG(cudaGetLastError());
    size_t tmp_bytes = 0;
    CKG(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_sorted, (int)total, 0, 64, st));
    CKG(cudaMalloc(&d_tmp, tmp_bytes));
    CKG(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_sorted, (int)total, 0, 64, st));
    cudaFree(d_tmp); d_tmp = nullptr;
    // d_keys is free again: reuse it for the unique keys
    d_unique = d_keys;
    CKG(cudaMalloc(&d_counts, total * 4));
    CKG(cudaMalloc(&d_nruns, sizeof(int)));
    tmp_bytes = 0;
    CKG(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, d_sorted, d_unique, d_counts, d_nruns, (int)total, st));
    CKG(cudaMalloc(&d_tmp, tmp_bytes));
    CKG(cub::DeviceRunLengthEncode::Encode(d_tmp, tmp_bytes, d_sorted, d_unique, d_counts, d_nruns, (int)total, st));
    CKG(cudaMalloc(&d_nsolid, 8));
    CKG(cudaMemsetAsync(d_nsolid, 0, 8, st));
    k_count_solid<<<grid_for(total, 256, sms), 256, 0, st>>>(d_counts, d_nruns, d_unique, (uint32_t)solid, d_nsolid);
    CKG(cudaGetLastError());
    unsigned long long n_solid = 0; int n_runs = 0;
    CKG(cudaMemcpyAsync(&n_solid, d_nsolid, 8, cudaMemcpyDeviceToHost, st));
    CKG(cudaMemcpyAsync(&n_runs, d_nruns, sizeof(int), cudaMemcpyDeviceToHost, st));
    CKG(cudaStreamSynchronize(st));
    CKG(table_alloc(dt, k, n_solid, sms, st));
    CKG(table_fill(dt, d_unique, d_counts, (uint32_t)solid, (uint64_t)n_runs, sms, st));
  }
done:
#undef CKG
  cudaFree(d_keys); cudaFree(d_sorted);
  cudaFree(d_counts); cudaFree(d_nruns); cudaFree(d_tmp); cudaFree(d_nsolid);
  return err;
}

cudaError_t table_from_reads(DeviceTable *dt, int k, const char *bases_host, const uint64_t *offsets_host, uint64_t n_reads,
                             int solid, int sms, cudaStream_t st) {
  // per-read k-mer counts -> exclusive offsets (host; linear in the number of reads)
  std::vector<uint64_t> koff(n_reads + 1);
  uint64_t total = 0;
  for (uint64_t i = 0; i < n_reads; i++) {
    koff[i] = total;
    const uint64_t len = offsets_host[i + 1] - offsets_host[i];
    total += len >= (uint64_t)k ? len - k + 1 : 0;
  }
  koff[n_reads] = total;
  const uint64_t nbytes = offsets_host[n_reads];
  char *d_bases = nullptr; uint64_t *d_off = nullptr, *d_koff = nullptr;
  cudaError_t err = cudaMalloc(&d_bases, nbytes + 1);
  if (err == cudaSuccess) err = cudaMalloc(&d_off, (n_reads + 1) * 8);
  if (err == cudaSuccess) err = cudaMalloc(&d_koff, (n_reads + 1) * 8);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_bases, bases_host, nbytes, cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_off, offsets_host, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_koff, koff.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess) err = table_from_device_reads(dt, k, d_bases, d_off, d_koff, n_reads, total, solid, sms, st);
  cudaFree(d_bases); cudaFree(d_off); cudaFree(d_koff);
  return err;
}

// ---------------------------------------------------------------------------------------------- reads parsed on the device
// The reads file(s) go to the device as raw bytes (FASTA, sequences over any number of lines, or four-line FASTQ) and
// are parsed there: a scan labels every byte with the kind of line it is on, the sequence characters are compacted into
// the concatenated-bases array the counting kernels read, and the read boundaries fall out of the same prefix sums.
// (The reference reads banks through GATB's BankFasta, gatb-core/src/gatb/bank/impl/BankFasta.cpp; only the sequences
// and their boundaries matter to the graph.)
struct LineState {   // per byte: kind of the line that holds it, and the number of line breaks before it
  uint32_t kind;     // 0 = unknown yet, 1 = FASTA sequence line, 2 = FASTA header line
  __host__ __device__ LineState() : kind(0) {}
  __host__ __device__ explicit LineState(uint32_t k) : kind(k) {}
};
struct LastKind {
  __host__ __device__ LineState operator()(const LineState &a, const LineState &b) const { return b.kind ? b : a; }
};

// FASTA: kind of the line each byte is on (scan of "a line starts here: header or not")
__global__ void k_fasta_line_starts(const char *bytes, uint64_t n, LineState *st) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const bool line_start = i == 0 || bytes[i - 1] == '\n';
    st[i] = LineState(line_start ? (bytes[i] == '>' ? 2u : 1u) : 0u);
  }
}
// keep[i]: byte i is a sequence character; head[i]: byte i starts a record (its header line)
__global__ void k_fasta_flags(const char *bytes, uint64_t n, const LineState *st, uint8_t *keep, uint8_t *head) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const char c = bytes[i];
    keep[i] = (st[i].kind == 1 && c != '\n' && c != '\r' && c != ' ' && c != '\t') ? 1 : 0;
    head[i] = (c == '>' && (i == 0 || bytes[i - 1] == '\n')) ? 1 : 0;
  }
}
// FASTQ (four lines per record): line number of every byte = line breaks before it
__global__ void k_newline_flags(const char *bytes, uint64_t n, uint32_t *nl) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) nl[i] = bytes[i] == '\n' ? 1u : 0u;
}
__global__ void k_fastq_flags(const char *bytes, uint64_t n, const uint32_t *line_no, uint8_t *keep, uint8_t *head) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const char c = bytes[i];
    const uint32_t q = line_no[i] & 3u;
    keep[i] = (q == 1 && c != '\n' && c != '\r') ? 1 : 0;
    head[i] = (q == 0 && (i == 0 || bytes[i - 1] == '\n') && c != '\n') ? 1 : 0;
  }
}
// offsets[r] = sequence characters before record r (kept_before at its header byte); bases compacted
__global__ void k_scatter_parsed(const char *bytes, uint64_t n, const uint8_t *keep, const uint8_t *head, const uint64_t *kept_before,
                                 const uint64_t *heads_before, char *bases, uint64_t *offsets) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    if (keep[i]) bases[kept_before[i]] = bytes[i];
    if (head[i]) offsets[heads_before[i]] = kept_before[i];
  }
}
__global__ void k_kmer_counts(const uint64_t *offsets, uint64_t n_reads, int k, uint64_t *cnt) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i <= n_reads; i += stride) {
    if (i == n_reads) { cnt[i] = 0; continue; }
    const uint64_t len = offsets[i + 1] - offsets[i];
    cnt[i] = len >= (uint64_t)k ? len - k + 1 : 0;
  }
}

struct U8ToU64 { __host__ __device__ uint64_t operator()(uint8_t v) const { return v; } };

cudaError_t table_from_read_files(DeviceTable *dt, int k, const char *bytes_host, uint64_t n, int solid, int sms, cudaStream_t st,
                                  uint64_t *n_reads_out) {
  if (n_reads_out) *n_reads_out = 0;
  // the first non-blank byte decides the format (BankFasta does the same)
  uint64_t p0 = 0;
  while (p0 < n && (bytes_host[p0] == '\n' || bytes_host[p0] == '\r' || bytes_host[p0] == ' ')) p0++;
  if (p0 == n) { cudaError_t e0 = table_alloc(dt, k, 0, sms, st); return e0 == cudaSuccess ? cudaStreamSynchronize(st) : e0; }
  const bool fastq = bytes_host[p0] == '@';
  if (!fastq && bytes_host[p0] != '>') return cudaErrorInvalidValue;
  bytes_host += p0; n -= p0;
  char *d_bytes = nullptr, *d_bases = nullptr;
  LineState *d_st = nullptr; uint32_t *d_nl = nullptr;
  uint8_t *d_keep = nullptr, *d_head = nullptr;
  uint64_t *d_kept_before = nullptr, *d_heads_before = nullptr, *d_off = nullptr, *d_koff = nullptr;
  void *d_tmp = nullptr;
  cudaError_t err = cudaSuccess;
#define CKG(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)
  {
    CKG(cudaMalloc(&d_bytes, n + 1));
    CKG(cudaMemcpyAsync(d_bytes, bytes_host, n, cudaMemcpyHostToDevice, st));
    CKG(cudaMalloc(&d_keep, n)); CKG(cudaMalloc(&d_head, n));
    size_t tb = 0;
    if (!fastq) {
      CKG(cudaMalloc(&d_st, n * sizeof(LineState)));
      k_fasta_line_starts<<<grid_for(n, 256, sms), 256, 0, st>>>(d_bytes, n, d_st);
      CKG(cub::DeviceScan::InclusiveScan(nullptr, tb, d_st, d_st, LastKind(), (int64_t)n, st));
      CKG(cudaMalloc(&d_tmp, tb));
      CKG(cub::DeviceScan::InclusiveScan(d_tmp, tb, d_st, d_st, LastKind(), (int64_t)n, st));
      cudaFree(d_tmp); d_tmp = nullptr;
      k_fasta_flags<<<grid_for(n, 256, sms), 256, 0, st>>>(d_bytes, n, d_st, d_keep, d_head);
      CKG(cudaGetLastError());
      cudaFree(d_st); d_st = nullptr;
    } else {
      CKG(cudaMalloc(&d_nl, n * 4));
      k_newline_flags<<<grid_for(n, 256, sms), 256, 0, st>>>(d_bytes, n, d_nl);
      CKG(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_nl, d_nl, (int64_t)n, st));
      CKG(cudaMalloc(&d_tmp, tb));
      CKG(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_nl, d_nl, (int64_t)n, st));
      cudaFree(d_tmp); d_tmp = nullptr;
      k_fastq_flags<<<grid_for(n, 256, sms), 256, 0, st>>>(d_bytes, n, d_nl, d_keep, d_head);
      CKG(cudaGetLastError());
      cudaFree(d_nl); d_nl = nullptr;
    }
    // prefix sums of the two flag arrays (64-bit: files beyond 4 GB)
    CKG(cudaMalloc(&d_kept_before, (n + 1) * 8));
    CKG(cudaMalloc(&d_heads_before, (n + 1) * 8));
    {
      thrust::transform_iterator<U8ToU64, const uint8_t *> kin((const uint8_t *)d_keep, U8ToU64()), hin((const uint8_t *)d_head, U8ToU64());
      tb = 0;
      CKG(cub::DeviceScan::ExclusiveSum(nullptr, tb, kin, d_kept_before, (int64_t)n + 1, st));
      CKG(cudaMalloc(&d_tmp, tb));
      CKG(cub::DeviceScan::ExclusiveSum(d_tmp, tb, kin, d_kept_before, (int64_t)n + 1, st));
      CKG(cub::DeviceScan::ExclusiveSum(d_tmp, tb, hin, d_heads_before, (int64_t)n + 1, st));
      cudaFree(d_tmp); d_tmp = nullptr;
    }
    uint64_t n_bases = 0, n_reads = 0;
    CKG(cudaMemcpyAsync(&n_bases, d_kept_before + n, 8, cudaMemcpyDeviceToHost, st));
    CKG(cudaMemcpyAsync(&n_reads, d_heads_before + n, 8, cudaMemcpyDeviceToHost, st));
    CKG(cudaStreamSynchronize(st));
    if (n_reads_out) *n_reads_out = n_reads;
    CKG(cudaMalloc(&d_bases, n_bases + 1));
    CKG(cudaMalloc(&d_off, (n_reads + 1) * 8));
    CKG(cudaMalloc(&d_koff, (n_reads + 2) * 8));
    k_scatter_parsed<<<grid_for(n, 256, sms), 256, 0, st>>>(d_bytes, n, d_keep, d_head, d_kept_before, d_heads_before, d_bases, d_off);
    CKG(cudaGetLastError());
    CKG(cudaMemcpyAsync(d_off + n_reads, &n_bases, 8, cudaMemcpyHostToDevice, st));
    CKG(cudaStreamSynchronize(st));   // n_bases lives on this stack frame
    cudaFree(d_bytes); d_bytes = nullptr; cudaFree(d_keep); d_keep = nullptr; cudaFree(d_head); d_head = nullptr;
    cudaFree(d_kept_before); d_kept_before = nullptr; cudaFree(d_heads_before); d_heads_before = nullptr;
    // k-mer occurrences per read -> exclusive offsets
    k_kmer_counts<<<grid_for(n_reads + 1, 256, sms), 256, 0, st>>>(d_off, n_reads, k, d_koff);
    CKG(cudaGetLastError());
    tb = 0;
    CKG(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_koff, d_koff, (int64_t)n_reads + 1, st));
    CKG(cudaMalloc(&d_tmp, tb));
    CKG(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_koff, d_koff, (int64_t)n_reads + 1, st));
    cudaFree(d_tmp); d_tmp = nullptr;
    uint64_t total = 0;
    CKG(cudaMemcpyAsync(&total, d_koff + n_reads, 8, cudaMemcpyDeviceToHost, st));
    CKG(cudaStreamSynchronize(st));
    CKG(table_from_device_reads(dt, k, d_bases, d_off, d_koff, n_reads, total, solid, sms, st));
  }
done:
#undef CKG
  cudaFree(d_bytes); cudaFree(d_bases); cudaFree(d_st); cudaFree(d_nl); cudaFree(d_keep); cudaFree(d_head);
  cudaFree(d_kept_before); cudaFree(d_heads_before); cudaFree(d_off); cudaFree(d_koff); cudaFree(d_tmp);
  return err;
}

// ---------------------------------------------------------------------------------------------- many small graphs
// Per-gap filtered-reads mode: the reference's wrapper runs one Gap2Seq-core process - one GATB
// graph construction - per gap on the reads that map near it (src/Gap2Seq.py:132-220, src/Gap2Seq.cpp:227-283).  Here
// the graphs of a whole batch of gaps are built in ONE device pass: every k-mer occurrence is tagged with its graph,
// one radix sort orders them by (graph, k-mer), an occurrence that starts a run of at least `solid` equal neighbours
// is a solid k-mer of its graph, and all tables live in one slot arena (one power-of-two region per graph).
__global__ void k_extract_tagged(const char *bases, const uint64_t *offsets, const uint64_t *kmer_offsets, uint64_t n_reads,
                                 const uint32_t *read_graph, int k, uint64_t *keys, uint32_t *tags, uint64_t total) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    uint64_t lo = 0, hi = n_reads;
    while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (kmer_offsets[mid] <= i) lo = mid; else hi = mid; }
    const char *s = bases + offsets[lo] + (i - kmer_offsets[lo]);
    uint64_t v = 0; uint32_t bad = 0;
    for (int j = 0; j < k; j++) {
      const unsigned char c = (unsigned char)s[j];
      bad |= (c >> 3) & 1;
      v = (v << 2) | ((c >> 1) & 3);
    }
    keys[i] = bad ? KEY_EMPTY : canonical(v, k).key;
    tags[i] = read_graph[lo];
  }
}

// occurrences sorted by (tag, key): position i starts a solid run iff it differs from i - 1 and equals i + solid - 1
__global__ void k_solid_heads(const uint64_t *keys, const uint32_t *tags, uint64_t total, uint32_t solid, uint64_t *out_keys,
                              uint32_t *out_tags, unsigned long long *cursor, uint32_t *per_graph) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const uint64_t key = keys[i];
    const uint32_t tag = tags[i];
    if (key == KEY_EMPTY) continue;
    if (i > 0 && keys[i - 1] == key && tags[i - 1] == tag) continue;
    const uint64_t j = i + (solid > 0 ? solid - 1 : 0);
    if (j >= total || keys[j] != key || tags[j] != tag) continue;
    const unsigned long long pos = atomicAdd(cursor, 1ULL);
    out_keys[pos] = key; out_tags[pos] = tag;
    atomicAdd(&per_graph[tag], 1u);
  }
}

__global__ void k_insert_multi(const Table *tables, const uint64_t *keys, const uint32_t *tags, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const Table t = tables[tags[i]];
    Slot *slots = const_cast<Slot *>(t.slots);
    const uint64_t key = keys[i];
    uint64_t h = mix64(key) & t.cap_mask;
    for (;;) {
      unsigned long long old = atomicCAS((unsigned long long *)&slots[h].key, (unsigned long long)KEY_EMPTY, (unsigned long long)key);
      if (old == KEY_EMPTY || old == key) break;
      h = (h + 1) & t.cap_mask;
    }
  }
}

// adjacency byte of every solid k-mer inside ITS graph; remembers where the k-mer sits for the hint pass
__global__ void k_adjacency_multi(const Table *tables, const uint64_t *keys, const uint32_t *tags, uint64_t n, uint64_t *slot_of) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const Table t = tables[tags[i]];
    const uint64_t key = keys[i];
    uint32_t adj = 0, dummy;
#pragma unroll
    for (uint32_t nt = 0; nt < 4; nt++) {
      if (table_find(t, canonical(succ_kmer(key, nt, t), t.k).key, &dummy) >= 0) adj |= 1u << nt;
      if (table_find(t, canonical(pred_kmer(key, nt, t), t.k).key, &dummy) >= 0) adj |= 1u << (4 + nt);
    }
    uint64_t h = mix64(key) & t.cap_mask;
    while (t.slots[h].key != key) h = (h + 1) & t.cap_mask;
    slot_of[i] = h;
    // the adjacency byte lives in the low byte of val and is only read (through the read-only path) by later launches
    const_cast<Slot *>(t.slots)[h].val = adj;
  }
}

__global__ void k_hints_multi(const Table *tables, const uint64_t *keys, const uint32_t *tags, uint64_t n, const uint64_t *slot_of, Slot *hints) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const Table t = tables[tags[i]];
    Slot o{keys[i], 0, {0, 0}};
    compute_hints(t, keys[i], (uint32_t)(t.slots[slot_of[i]].val & 0xFF), &o);
    hints[i] = o;
  }
}

__global__ void k_apply_hints_multi(const Table *tables, const uint32_t *tags, uint64_t n, const uint64_t *slot_of, const Slot *hints) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    Slot *s = const_cast<Slot *>(tables[tags[i]].slots) + slot_of[i];
    s->val = hints[i].val; s->h[0] = hints[i].h[0]; s->h[1] = hints[i].h[1];
  }
}

void multi_table_release(MultiTable *mt) {
  if (mt->slots) cudaFree(mt->slots);
  if (mt->d_tables) cudaFree(mt->d_tables);
  mt->slots = nullptr; mt->d_tables = nullptr; mt->total_slots = 0; mt->n_kmers = 0; mt->n_graphs = 0;
  mt->h_tables.clear(); mt->h_sizes.clear();
}

cudaError_t tables_from_reads(MultiTable *mt, int k, const char *bases_host, const uint64_t *offsets_host, uint64_t n_reads,
                              const uint64_t *graph_first_read, uint64_t n_graphs, int solid, int sms, cudaStream_t st) {
  multi_table_release(mt);
  if (n_graphs == 0 || n_graphs >= (1ULL << 31)) return cudaErrorInvalidValue;
  std::vector<uint64_t> koff(n_reads + 1);
  std::vector<uint32_t> rgraph(std::max<uint64_t>(n_reads, 1));
  uint64_t total = 0;
  for (uint64_t i = 0; i < n_reads; i++) {
    koff[i] = total;
    const uint64_t len = offsets_host[i + 1] - offsets_host[i];
    total += len >= (uint64_t)k ? len - k + 1 : 0;
  }
  koff[n_reads] = total;
  for (uint64_t g = 0; g < n_graphs; g++) {
    if (graph_first_read[g] > graph_first_read[g + 1] || graph_first_read[g + 1] > n_reads) return cudaErrorInvalidValue;
    for (uint64_t r = graph_first_read[g]; r < graph_first_read[g + 1]; r++) rgraph[r] = (uint32_t)g;
  }
  if (graph_first_read[0] != 0 || graph_first_read[n_graphs] != n_reads) return cudaErrorInvalidValue;
  if (total >= (1ULL << 31)) return cudaErrorInvalidValue;   // one sort pass; larger batches are the caller's to split
  const uint64_t nbytes = offsets_host[n_reads];
  char *d_bases = nullptr; uint64_t *d_off = nullptr, *d_koff = nullptr, *d_keys = nullptr, *d_keys2 = nullptr, *d_slot_of = nullptr;
  uint32_t *d_rgraph = nullptr, *d_tags = nullptr, *d_tags2 = nullptr, *d_per_graph = nullptr;
  unsigned long long *d_cursor = nullptr; void *d_tmp = nullptr; Slot *d_hints = nullptr;
  std::vector<uint32_t> sizes(n_graphs, 0);
  cudaError_t err = cudaSuccess;
  int tag_bits = 1;
  while ((1ULL << tag_bits) < n_graphs) tag_bits++;
#define CKG(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)
  {
    unsigned long long n_solid = 0;
    CKG(cudaMalloc(&d_per_graph, n_graphs * 4));
    CKG(cudaMemsetAsync(d_per_graph, 0, n_graphs * 4, st));
    CKG(cudaMalloc(&d_cursor, 8));
    CKG(cudaMemsetAsync(d_cursor, 0, 8, st));
    if (total > 0) {
      CKG(cudaMalloc(&d_bases, nbytes + 1));
      CKG(cudaMalloc(&d_off, (n_reads + 1) * 8));
      CKG(cudaMalloc(&d_koff, (n_reads + 1) * 8));
      CKG(cudaMalloc(&d_rgraph, n_reads * 4));
      CKG(cudaMemcpyAsync(d_bases, bases_host, nbytes, cudaMemcpyHostToDevice, st));
      CKG(cudaMemcpyAsync(d_off, offsets_host, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
      CKG(cudaMemcpyAsync(d_koff, koff.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
      CKG(cudaMemcpyAsync(d_rgraph, rgraph.data(), n_reads * 4, cudaMemcpyHostToDevice, st));
      CKG(cudaMalloc(&d_keys, total * 8)); CKG(cudaMalloc(&d_keys2, total * 8));
      CKG(cudaMalloc(&d_tags, total * 4)); CKG(cudaMalloc(&d_tags2, total * 4));
      k_extract_tagged<<<grid_for(total, 256, sms), 256, 0, st>>>(d_bases, d_off, d_koff, n_reads, d_rgraph, k, d_keys, d_tags, total);
      CKG(cudaGetLastError());
      // least significant digit first: by k-mer, then (stable) by graph
      size_t tb = 0, tb2 = 0;
      CKG(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_keys, d_keys2, d_tags, d_tags2, (int)total, 0, 64, st));
      CKG(cub::DeviceRadixSort::SortPairs(nullptr, tb2, d_tags2, d_tags, d_keys2, d_keys, (int)total, 0, tag_bits, st));
      CKG(cudaMalloc(&d_tmp, std::max(tb, tb2)));
      CKG(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_keys, d_keys2, d_tags, d_tags2, (int)total, 0, 64, st));
      CKG(cub::DeviceRadixSort::SortPairs(d_tmp, tb2, d_tags2, d_tags, d_keys2, d_keys, (int)total, 0, tag_bits, st));
      // (d_keys, d_tags) are sorted by (graph, k-mer); the solid ones are appended to (d_keys2, d_tags2)
      k_solid_heads<<<grid_for(total, 256, sms), 256, 0, st>>>(d_keys, d_tags, total, (uint32_t)std::max(solid, 1), d_keys2, d_tags2,
                                                               d_cursor, d_per_graph);
      CKG(cudaGetLastError());
    }
    CKG(cudaMemcpyAsync(&n_solid, d_cursor, 8, cudaMemcpyDeviceToHost, st));
    CKG(cudaMemcpyAsync(sizes.data(), d_per_graph, n_graphs * 4, cudaMemcpyDeviceToHost, st));
    CKG(cudaStreamSynchronize(st));
    // geometry: one power-of-two region per graph, load <= 0.5 like the single table
    mt->h_tables.resize(n_graphs);
    uint64_t slots_total = 0;
    for (uint64_t g = 0; g < n_graphs; g++) {
      uint64_t cap = 16;
      while (cap < 2ULL * sizes[g] + 2) cap <<= 1;
      Table &t = mt->h_tables[g];
      t.slots = (const Slot *)(uintptr_t)slots_total;   // offset for now
      t.cap_mask = cap - 1; t.k = k; t.kmask = k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
      slots_total += cap;
    }
    CKG(cudaMalloc(&mt->slots, slots_total * sizeof(Slot)));
    for (uint64_t g = 0; g < n_graphs; g++) mt->h_tables[g].slots = mt->slots + (uintptr_t)mt->h_tables[g].slots;
    CKG(cudaMalloc(&mt->d_tables, n_graphs * sizeof(Table)));
    CKG(cudaMemcpyAsync(mt->d_tables, mt->h_tables.data(), n_graphs * sizeof(Table), cudaMemcpyHostToDevice, st));
    k_clear_slots<<<grid_for(slots_total, 256, sms), 256, 0, st>>>(mt->slots, slots_total);
    CKG(cudaGetLastError());
    if (n_solid > 0) {
      CKG(cudaMalloc(&d_slot_of, n_solid * 8));
      CKG(cudaMalloc(&d_hints, n_solid * sizeof(Slot)));
      k_insert_multi<<<grid_for(n_solid, 256, sms), 256, 0, st>>>(mt->d_tables, d_keys2, d_tags2, n_solid);
      k_adjacency_multi<<<grid_for(n_solid, 256, sms), 256, 0, st>>>(mt->d_tables, d_keys2, d_tags2, n_solid, d_slot_of);
      k_hints_multi<<<grid_for(n_solid, 256, sms), 256, 0, st>>>(mt->d_tables, d_keys2, d_tags2, n_solid, d_slot_of, d_hints);
      k_apply_hints_multi<<<grid_for(n_solid, 256, sms), 256, 0, st>>>(mt->d_tables, d_tags2, n_solid, d_slot_of, d_hints);
      CKG(cudaGetLastError());
    }
    CKG(cudaStreamSynchronize(st));
    mt->total_slots = slots_total; mt->n_kmers = n_solid; mt->n_graphs = n_graphs;
    mt->h_sizes.assign(sizes.begin(), sizes.end());
  }
done:
#undef CKG
  cudaFree(d_bases); cudaFree(d_off); cudaFree(d_koff); cudaFree(d_rgraph); cudaFree(d_keys); cudaFree(d_keys2);
  cudaFree(d_tags); cudaFree(d_tags2); cudaFree(d_per_graph); cudaFree(d_cursor); cudaFree(d_tmp); cudaFree(d_slot_of); cudaFree(d_hints);
  if (err != cudaSuccess) multi_table_release(mt);
  return err;
}

cudaError_t table_export(const DeviceTable &dt, uint64_t *keys_host, uint8_t *adj_host, int sms, cudaStream_t st) {
  if (dt.n == 0) return cudaSuccess;
  uint64_t *d_keys = nullptr, *d_sorted = nullptr; unsigned long long *d_cur = nullptr; void *d_tmp = nullptr;
  uint8_t *d_adj = nullptr;
  cudaError_t err = cudaSuccess;
#define CKG(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)
  {
    CKG(cudaMalloc(&d_keys, dt.n * 8));
    CKG(cudaMalloc(&d_sorted, dt.n * 8));
    CKG(cudaMalloc(&d_cur, 8));
    CKG(cudaMemsetAsync(d_cur, 0, 8, st));
    k_compact<<<grid_for(dt.cap, 256, sms), 256, 0, st>>>(dt.slots, dt.cap, d_keys, d_cur);
    CKG(cudaGetLastError());
    size_t tmp_bytes = 0;
    CKG(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_sorted, (int)dt.n, 0, 64, st));
    CKG(cudaMalloc(&d_tmp, tmp_bytes));
    CKG(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_sorted, (int)dt.n, 0, 64, st));
    CKG(cudaMemcpyAsync(keys_host, d_sorted, dt.n * 8, cudaMemcpyDeviceToHost, st));
    if (adj_host) {
      CKG(cudaMalloc(&d_adj, dt.n));
      k_lookup_adj<<<(unsigned)((dt.n + 255) / 256), 256, 0, st>>>(table_view(dt), d_sorted, dt.n, d_adj);
      CKG(cudaGetLastError());
      CKG(cudaMemcpyAsync(adj_host, d_adj, dt.n, cudaMemcpyDeviceToHost, st));
    }
    CKG(cudaStreamSynchronize(st));
  }
done:
#undef CKG
  cudaFree(d_keys); cudaFree(d_sorted); cudaFree(d_cur); cudaFree(d_tmp); cudaFree(d_adj);
  return err;
}

cudaError_t table_contains(const DeviceTable &dt, const char *ascii_host, uint64_t n, uint8_t *out_host, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  char *d_a = nullptr; uint8_t *d_o = nullptr;
  cudaError_t err = cudaSuccess;
#define CKG(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)
  {
    CKG(cudaMalloc(&d_a, n * dt.k));
    CKG(cudaMalloc(&d_o, n));
    CKG(cudaMemcpyAsync(d_a, ascii_host, n * dt.k, cudaMemcpyHostToDevice, st));
    k_contains<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(table_view(dt), d_a, n, d_o);
    CKG(cudaGetLastError());
    CKG(cudaMemcpyAsync(out_host, d_o, n, cudaMemcpyDeviceToHost, st));
    CKG(cudaStreamSynchronize(st));
  }
done:
#undef CKG
  cudaFree(d_a); cudaFree(d_o);
  return err;
}

}  // namespace g2s
