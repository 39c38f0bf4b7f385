/* g2s.h - C ABI of the B200 gap-filling engine (libg2s_gpu.so).
 *
 * The reference (rikuu/Gap2Seq) has no plugin/FFI layer: its hot path is the member function
 *     int Gap2Seq::fill_gap(const Graph&, const std::string &kmer_left, const std::string &kmer_right,
 *                           int gap_len, int k, int gap_err, int left_max_fuz, int right_max_fuz,
 *                           int *left_fuz, int *right_fuz, long long max_mem, char *fill,
 *                           bool skip_confident, bool all_paths, struct subgraph_stats *substats)
 * declared at src/Gap2Seq.hpp:74-76, defined at src/Gap2Seq.cpp:858-1556 and called from exactly two
 * places, src/Gap2Seq.cpp:252 (single-gap mode) and src/Gap2Seq.cpp:380 (scaffold mode).  The entry
 * points below are what a cgo/JNI/ctypes/C++ binding for that path binds; INTEGRATION.md shows the
 * reference-side change.  Plain pointers and sizes only; all buffers are caller owned.
 *
 * Every function returns 0 (or a non-negative count where stated) on success and a negative
 * G2S_E_* code on failure; g2s_last_error() then describes the failure.  There is no CPU fallback:
 * without a CUDA device every compute call fails with G2S_E_CUDA.
 */
#ifndef G2S_H
#define G2S_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2S_ABI_VERSION 4

enum {
  G2S_OK = 0,
  G2S_E_CUDA = -1,      /* CUDA runtime error or no device */
  G2S_E_ARG = -2,       /* invalid argument */
  G2S_E_STATE = -3,     /* call out of order (e.g. fill before a graph was imported) */
  G2S_E_NOMEM = -4      /* host or device allocation failed */
};

/* per-gap status bits (g2s_result.status) */
enum {
  G2S_ST_ORDER_DEPENDENT = 1,  /* a border set received both strands of one k-mer: the reference's result depends on
                                  libstdc++ hashtable iteration order; reported, not reproduced */
  G2S_ST_MEM_LIMIT = 2,        /* the DP outgrew the largest device tier (the analogue of the reference's
                                  "Memory limit exceeded", src/Gap2Seq.cpp:147); count is -1 */
  G2S_ST_GENERAL_SUBGRAPH = 4, /* informational: the path subgraph had a k-mer at several levels, SCC analysis ran */
  G2S_ST_UNABLE_TO_BACKTRACE = 8,
  G2S_ST_BIG_TIER = 16,        /* informational: the gap was re-run in the wide (one CTA per gap) tier */
  G2S_ST_DEVICE_TRACED = 32,   /* informational: the traceback had no choice to make (one trellis cell per level, one path
                                  length), the device wrote the fill itself and rand() was only advanced */
  G2S_ST_PHANTOM_FLANK = 64    /* the caller's graph (g2s_gap.present_bits: GATB's contains() per flank k-mer, Bloom AND NOT
                                  cFP AND MPHF, Graph.hpp:1249-1272) says a flank k-mer exists that is NOT in the imported
                                  solid set - a Bloom false positive the reference would start a walk from - or
                                  the other way round.  The gap is filled on the exact set; the reference may differ */
};

/* gap flags */
enum {
  G2S_GAP_SKIP_CONFIDENT = 1,  /* -all-upper: skip the safe/unsafe classification (fill_gap's skip_confident) */
  G2S_GAP_ALL_PATHS = 2,       /* default mode; cleared by -best-only (fill_gap's all_paths) */
  G2S_GAP_PRESENT_BITS = 4     /* g2s_gap.present_bits is valid */
};

typedef struct g2s_engine g2s_engine;

typedef struct g2s_config {
  int32_t abi_version;   /* G2S_ABI_VERSION */
  int32_t device;        /* CUDA device ordinal */
  int32_t k;             /* k-mer size, odd, 3 <= k <= 31 (the -k flag, src/Gap2Seq.cpp:164) */
  int32_t max_fuz;       /* upper bound of left/right_max_fuz over all gaps, <= 31 (the -fuz flag) */
  int64_t max_mem_bytes; /* device memory budget for DP scratch; 0 = 60 % of the free memory (the -max-mem flag) */
  int32_t host_threads;  /* threads for the host-side subgraph analysis; 0 = hardware concurrency (-nb-cores) */
  int32_t reserved;
} g2s_config;

/* mirror of struct subgraph_stats, src/Gap2Seq.hpp:51-59 */
typedef struct g2s_subgraph_stats {
  uint64_t vertices, edges, nontrivial_components, size_nontrivial_components, vertices_final, edges_final;
} g2s_subgraph_stats;

/* one call of fill_gap: the arguments of src/Gap2Seq.hpp:74-76 that vary per gap */
typedef struct g2s_gap {
  const char *kmer_left;    /* k + left_max_fuz characters (need not be NUL terminated) */
  const char *kmer_right;   /* k + right_max_fuz characters */
  int32_t gap_len, gap_err, left_max_fuz, right_max_fuz;
  int32_t flags;            /* G2S_GAP_* */
  int32_t graph;            /* 0, or the gap's graph after g2s_build_graphs_from_reads (per-gap filtered-reads mode) */
  const char *kmer_right_starts; /* NULL (= kmer_right) or k + right_max_fuz characters whose k-mers start the right
                               pass.  fill_gap takes those from the END of its kmer_right string but the targets from its
                               beginning (src/Gap2Seq.cpp:871-885 vs :1119); they differ only when the caller's right flank
                               is longer than k + right_max_fuz, which single-gap mode allows (src/Gap2Seq.cpp:239-252) */
  uint64_t present_bits;    /* with G2S_GAP_PRESENT_BITS: what the CALLER's graph answers to contains() for the flank k-mers
                               fill_gap looks up (src/Gap2Seq.cpp:884, 957, 1000, 1086): bit d (0..left_max_fuz) = left
                               k-mer at offset d, bit 32 + j (0..right_max_fuz) = right k-mer at offset j.  The engine
                               compares them with its exact table and reports G2S_ST_PHANTOM_FLANK on disagreement */
} g2s_gap;

typedef struct g2s_result {
  int32_t count;            /* fill_gap's return value: paths found, 0, or -1 (G2S_ST_MEM_LIMIT) */
  int32_t left_fuz, right_fuz;
  int32_t status;           /* G2S_ST_* bits */
  int32_t fill_len;         /* strlen(&fill[left_max_fuz - left_fuz]); ends with the right k-mer; 0 if not traced */
  int32_t substats_valid;   /* substats were written (Phase B ran) */
  int32_t forward_count;    /* count after the first-found check (before the all-paths recount) */
  int32_t n_path_lengths, path_lengths[2], reached_fuz;
  int32_t rand_calls;       /* libc rand() calls consumed by this gap's traceback */
  g2s_subgraph_stats substats;
  uint64_t cells_right, cells_left;   /* DP cells materialised (== the reference's -DSTATS "Nonzeros") */
  uint64_t nodes;                     /* distinct canonical k-mers touched by the gap */
} g2s_result;

typedef struct g2s_timing {
  double h2d_ms, kernel_ms, d2h_ms, host_ms;   /* of the last g2s_fill_batch / g2s_run_resident call */
  uint64_t h2d_bytes, d2h_bytes;
  uint32_t launches;                            /* kernels launched by that call */
  uint32_t gaps_big_tier;
  /* device-memory budget (the -max-mem analogue): what the tier ladder did in that call */
  uint32_t max_tier_run;                        /* highest tier a pass ran in (0 = thin tier only) */
  uint32_t gaps_rerun_scratch;                  /* gap runs that ended because the scratch slice of their tier was full */
  uint32_t gaps_rerun_out_full;                 /* gap runs that ended because the pass's pack arena was full */
  uint32_t gaps_mem_limit;                      /* gaps given up: no larger tier fits g2s_config.max_mem_bytes (count -1) */
  /* hash work of the gap runs that produced the results (re-runs in a larger tier count once) */
  uint64_t table_lookups;                       /* hash-gathers in the global k-mer table (one 32-byte sector each when the
                                                   first probe hits): flank k-mers, scout segments, expanded successors */
  uint64_t node_ops;                            /* find / find-or-insert operations on the per-gap node tables */
  /* which kernels ran: the thin tier (k_fill_thin<H,M>, thin_shape 0/1/2 = <1,1>/<2,4>/<4,8> warps per gap) and the
     wide tiers (k_fill_cta, one CTA per gap) */
  double kernel_ms_thin, kernel_ms_wide;
  uint32_t launches_thin, launches_wide;
  uint32_t thin_shape, wide_cluster;            /* wide tiers: CTAs per gap of the last pass (1: k_fill_cta, 2-8: k_fill_cluster) */
  uint64_t cells_thin, cells_wide;              /* DP cells (right + forward) of the gaps each of them finished */
  /* the chain tier (k_chain_walk + k_chain_finish): gaps whose borders never exceed one state are finished there
     without per-gap scratch tables; every other gap goes on to the thin tier */
  double kernel_ms_chain;
  uint32_t launches_chain, gaps_chain;
  uint64_t cells_chain;
} g2s_timing;

const char *g2s_last_error(void);
int g2s_abi_version(void);

int g2s_create(const g2s_config *cfg, g2s_engine **out);
void g2s_destroy(g2s_engine *e);

/* Several GPUs of one box behind ONE engine (cfg->device is ignored, devices[0] is the primary GPU).  Gaps are
 * independent units (the reference hands records to -nb-cores Dispatcher workers, src/Gap2Seq.cpp:296-323, 426-432):
 * every wave of gaps of g2s_fill_batch / g2s_submit / g2s_fill_records is split longest-processing-time-first by
 * estimated cost over the GPUs, each GPU runs its shard on its own host thread, and the per-gap results are gathered in
 * input order so that ONE libc rand() stream drives the tracebacks - the output is byte-identical to a one-GPU run.
 * The k-mer table is built / imported on devices[0] and replicated to the others once (ncclBroadcast over NVLink; NCCL
 * is bound at run time).  The resident split API (g2s_upload_gaps ...) and g2s_table_raw/adopt address devices[0] only. */
int g2s_create_multi(const g2s_config *cfg, const int32_t *devices, int32_t n_devices, g2s_engine **out);
int g2s_num_devices(const g2s_engine *e);

/* Graph import: replaces the GATB neighbour queries (Graph::successors/predecessors/contains,
 * gatb-core/src/gatb/debruijn/impl/Graph.cpp:1284-1345, 2193-2213) by an open-addressing table of 2-bit
 * packed canonical k-mers (GATB code A=0,C=1,T=2,G=3) with an 8-bit in/out edge mask per node
 * (the byte Graph::precomputeAdjacency defines, Graph.cpp:3486-3530).  `canon` is a HOST array of the
 * solid canonical k-mers, e.g. the values of graph.iterator() (Graph.cpp:2243-2362). */
int g2s_import_kmers(g2s_engine *e, const uint64_t *canon, size_t n);

/* The same from the reads files themselves (FASTA with sequences over any number of lines, or four-line FASTQ; all files
 * of one call in the same format): the raw bytes go to the device and are parsed there (line kinds by a scan, sequence
 * characters compacted, read boundaries from the same prefix sums).  *n_reads_out (may be NULL) = records found. */
int g2s_build_from_read_files(g2s_engine *e, const char *const *paths, size_t n_paths, int solid, uint64_t *n_reads_out);

/* Graph import from a GATB graph file (<reads>.h5), the file Gap2Seq::execute loads instead of building the graph when
 * it is readable (src/Gap2Seq.cpp:171, 195-197): the partitions of group dsk/solid are read straight from the HDF5
 * container (gatb-core/src/gatb/debruijn/impl/Graph.cpp:171-184, 2243-2362 own the format) - no GATB, no libhdf5.
 * g2s_h5_read_solid is host-only: *canon_out (malloc'ed, release with g2s_free) receives the n canonical k-mers,
 * *kmer_size_out the file's "kmer_size" property (0 when absent).  g2s_import_h5 = read + g2s_import_kmers; it refuses
 * a file whose k-mer size differs from the engine's. */
int g2s_h5_read_solid(const char *path, uint64_t **canon_out, size_t *n_out, int32_t *kmer_size_out);
int g2s_import_h5(g2s_engine *e, const char *path);

/* Graph construction on the device, the stand-in for Graph::create(... -abundance-min solid ...)
 * (src/Gap2Seq.cpp:213-214): counts canonical k-mers of `n_reads` reads (ASCII, read i is
 * bases[offsets[i] .. offsets[i+1]) ) and keeps those seen >= solid times.  Timed separately from the fill. */
int g2s_build_from_reads(g2s_engine *e, const char *bases, const uint64_t *offsets, size_t n_reads, int solid);

/* Per-gap filtered-reads mode.  The reference's wrapper fills every gap in its own Gap2Seq-core process
 * on the reads that map near the gap - one GATB graph construction per gap (src/Gap2Seq.py:132-220, single-gap mode
 * src/Gap2Seq.cpp:227-283).  Here the graphs of a whole batch are built in ONE device pass: graph g is counted from reads
 * graph_first_read[g] .. graph_first_read[g + 1] - 1 of the concatenated reads (the ranges tile the reads in order), with
 * its own solidity filter, into its own region of one table arena.  Afterwards every gap of g2s_fill_batch / g2s_submit
 * names its graph in g2s_gap.graph and all gaps are filled in one pass, each against its own graph.  One-GPU engines;
 * replaces the single graph of g2s_import_kmers / g2s_build_from_reads (and is replaced by them). */
int g2s_build_graphs_from_reads(g2s_engine *e, const char *bases, const uint64_t *offsets, size_t n_reads,
                                const uint64_t *graph_first_read, size_t n_graphs, int solid);
int64_t g2s_graph_num_kmers(const g2s_engine *e, size_t graph);   /* solid k-mers of one of those graphs, -1 = no such graph */

/* Replication of a built table to other GPUs (one NCCL broadcast over NVLink, done by the caller - e.g.
 * torch.distributed.broadcast in bench.py): the owner exposes its device buffer, the receivers allocate an
 * identically shaped one, the caller copies/broadcasts the bytes device to device and then commits. */
int g2s_table_raw(g2s_engine *e, void **device_ptr, uint64_t *capacity_slots, uint64_t *n_kmers, uint64_t *bytes);
int g2s_table_adopt(g2s_engine *e, uint64_t capacity_slots, uint64_t n_kmers, void **device_ptr);

int64_t g2s_num_kmers(const g2s_engine *e);
/* copies the table contents (sorted canonical k-mers and their adjacency bytes) to host arrays of g2s_num_kmers() */
int g2s_export_kmers(g2s_engine *e, uint64_t *canon_sorted, uint8_t *adjacency);
/* Graph::contains for n canonical-or-not k-mers given as ASCII strings of length k, one per k bytes */
int g2s_contains(g2s_engine *e, const char *kmers_ascii, size_t n, uint8_t *out);

/* The drop-in for Gap2Seq::fill_gap (src/Gap2Seq.hpp:74-76): same arguments minus the Graph (owned by the
 * engine), same return value, same `fill` buffer contract (caller allocates gap_len+k+gap_err+lmf+rmf+3 bytes,
 * the valid string starts at &fill[left_max_fuz - *left_fuz]), consumes libc rand() exactly like the reference. */
int g2s_fill_gap(g2s_engine *e, const char *kmer_left, const char *kmer_right, int gap_len, int k, int gap_err,
                 int left_max_fuz, int right_max_fuz, int *left_fuz, int *right_fuz, long long max_mem, char *fill,
                 int skip_confident, int all_paths, g2s_subgraph_stats *substats);

/* g2s_fill_gap for a caller that keeps its own graph (INTEGRATION.md form 2): additionally takes the caller's
 * contains() answers for the flank k-mers (g2s_gap.present_bits) and returns the G2S_ST_* bits in *status. */
int g2s_fill_gap_ex(g2s_engine *e, const char *kmer_left, const char *kmer_right, int gap_len, int k, int gap_err,
                    int left_max_fuz, int right_max_fuz, int *left_fuz, int *right_fuz, long long max_mem, char *fill,
                    int skip_confident, int all_paths, g2s_subgraph_stats *substats, uint64_t present_bits,
                    int have_present_bits, int32_t *status);

/* Batched form: n independent gaps in one device pass; the host-side random traceback then replays libc
 * rand() in array order, so the results equal n successive g2s_fill_gap calls.  fill_out[i] (may be NULL)
 * receives the NUL-terminated fill string, i.e. what the reference leaves at &fill[left_max_fuz - left_fuz];
 * it must hold gap_len+k+gap_err+lmf+rmf+3 bytes. */
int g2s_fill_batch(g2s_engine *e, const g2s_gap *gaps, size_t n, g2s_result *results, char *const *fill_out);

/* Split form used by bench.py to time the device pass with inputs already resident in HBM:
 * upload once, run the device pass any number of times, then finish (D2H + host traceback) once. */
int g2s_upload_gaps(g2s_engine *e, const g2s_gap *gaps, size_t n);
int g2s_run_resident(g2s_engine *e);          /* device pass only; synchronous */
int g2s_finish_resident(g2s_engine *e, g2s_result *results, char *const *fill_out);

/* Detached batches: the device pass of n gaps without the traceback, so that several batches can be traced
 * later in a caller-chosen order (the scaffold driver needs this: a gap's flanks may depend on the previous
 * gap's count / right_fuz, which are known before the traceback, src/Gap2Seq.cpp:392-402). */
typedef struct g2s_batch g2s_batch;
int g2s_submit(g2s_engine *e, const g2s_gap *gaps, size_t n, g2s_batch **out);   /* consumes no rand() */
int g2s_batch_peek(const g2s_batch *b, size_t i, int32_t *count, int32_t *right_fuz);
int g2s_batch_trace(g2s_batch *b, size_t i, g2s_result *result, char *fill_out);  /* consumes rand() */
void g2s_batch_free(g2s_batch *b);

/* The gap driver of Gap2Seq::execute() (src/Gap2Seq.cpp:285-434): scans n_records scaffold records for N/n runs,
 * fills them through the engine in large batches, and returns for every record the filled sequence
 * (filled_out[i], malloc'ed, release with g2s_free) plus the stdout lines the reference prints per gap
 * (*log_out, malloc'ed).  Byte-identical to `Gap2Seq-core -nb-cores 1` for the same libc rand() state. */
typedef struct g2s_run_options {
  int32_t k, dist_error, fuz;
  int32_t skip_confident, unique_paths, all_paths;
} g2s_run_options;
int g2s_fill_records(g2s_engine *e, const char *const *seqs, const char *const *comments, size_t n_records,
                     const g2s_run_options *opt, char **filled_out, char **log_out, int32_t *gapcount,
                     int32_t *filledcount, g2s_result **gap_results, size_t *n_gap_results);
void g2s_free(void *p);

int g2s_get_timing(const g2s_engine *e, g2s_timing *out);
/* profiling aid: SM cycles per phase of the last device pass, summed over gaps (sum[14]) and of the slowest gap
 * (slowest[14]); [0..5]: init, right pass, forward pass, backward marking, pack, total; [6..10] forward detail: scout,
 * node phase, layout, found scan, level-by-level path; [11..13] counts: narrow batches, narrow levels, general levels */
int g2s_get_phase_cycles(const g2s_engine *e, double *sum, double *slowest);

#ifdef __cplusplus
}
#endif
#endif /* G2S_H */
