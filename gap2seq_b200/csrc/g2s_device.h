// g2s_device.h - host-callable entry points of the CUDA translation units (internal to libg2s_gpu.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "g2s_common.cuh"

namespace g2s {

struct DeviceTable {
  Slot *slots = nullptr;
  uint64_t cap = 0;   // slots (power of two)
  uint64_t n = 0;     // k-mers stored
  int k = 0;
};

Table table_view(const DeviceTable &dt);
cudaError_t table_alloc(DeviceTable *dt, int k, uint64_t n_keys, int sms, cudaStream_t st);
cudaError_t table_fill(DeviceTable *dt, const uint64_t *keys_dev, const uint32_t *counts_dev, uint32_t solid, uint64_t n,
                       int sms, cudaStream_t st);
cudaError_t table_from_reads(DeviceTable *dt, int k, const char *bases_host, const uint64_t *offsets_host,
                             uint64_t n_reads, int solid, int sms, cudaStream_t st);
// reads as the raw bytes of FASTA / four-line FASTQ text (several files: concatenated, each ending with a line break),
// parsed on the device; *n_reads_out = records found
cudaError_t table_from_read_files(DeviceTable *dt, int k, const char *bytes_host, uint64_t n, int solid, int sms, cudaStream_t st,
                                  uint64_t *n_reads_out);
cudaError_t table_export(const DeviceTable &dt, uint64_t *keys_host, uint8_t *adj_host, int sms, cudaStream_t st);
cudaError_t table_contains(const DeviceTable &dt, const char *ascii_host, uint64_t n, uint8_t *out_host, cudaStream_t st);

// many small graphs in one slot arena (per-gap filtered-reads mode, g2s_table.cu tables_from_reads)
struct MultiTable {
  Slot *slots = nullptr;          // the arena
  Table *d_tables = nullptr;      // device array, one view per graph
  std::vector<Table> h_tables;
  std::vector<uint32_t> h_sizes;  // solid k-mers per graph
  uint64_t total_slots = 0, n_kmers = 0, n_graphs = 0;
};
void multi_table_release(MultiTable *mt);
// graph g is built from reads graph_first_read[g] .. graph_first_read[g + 1] - 1 of the concatenated reads
cudaError_t tables_from_reads(MultiTable *mt, int k, const char *bases_host, const uint64_t *offsets_host, uint64_t n_reads,
                              const uint64_t *graph_first_read, uint64_t n_graphs, int solid, int sms, cudaStream_t st);

struct FillParams;
struct ChainParams;
// chain tier (g2s_chain.cuh): k_chain_walk + k_chain_finish over the gaps of p.order
cudaError_t launch_chain(const ChainParams &p, cudaStream_t st);
// thin tier: one CTA per gap, persistent over a device work queue; `shape` selects how many warps a gap gets
// (0: 1 helper + 1 main, 1: 2 + 4, 2: 4 + 8).  groups = number of scratch slices.
cudaError_t launch_fill_thin(const FillParams &p, int groups, int shape, cudaStream_t st);
// wide tier: one CTA (512 threads) per gap
cudaError_t launch_fill_cta(const FillParams &p, int groups, cudaStream_t st);
// wide tier with a thread-block cluster of `cluster` (2..8) CTAs per gap; groups = number of clusters = scratch slices
cudaError_t launch_fill_cluster(const FillParams &p, int groups, int cluster, cudaStream_t st);
int fill_max_resident_clusters(int cluster);
int fill_thin_shapes();
int fill_thin_threads(int shape);
int fill_max_resident_thin_groups(int sms, int shape);   // occupancy-derived number of thin groups that fit the device
int fill_max_resident_cta_groups(int sms);

}  // namespace g2s
