// g2s_fill.cuh - one gap of Gap2Seq::fill_gap (reference src/Gap2Seq.cpp:858-1556) as a level-synchronous
// frontier expansion executed cooperatively by one thread group (a warp for thin gaps, a whole CTA for
// wide ones).  Everything the reference keeps in unordered_map<Node, map_element> rows lives in a per-group
// scratch slice in HBM:
//
//   node table   open addressing, 32 B entries keyed by canonical k-mer (Node identity ignores the strand,
//                Graph.hpp:154-160): right-set membership, per-level stamps that de-duplicate a frontier,
//                the index of the node's cell on the current level (one per strand, like map_element's
//                forward / reverse vectors, Gap2Seq.cpp:492-493).
//   cells        structure of arrays, appended level by level; cells [lstart[c], lstart[c+1]) ARE the border of
//                level c, so the expansion reads its frontier coalesced.  A cell stores its oriented k-mer, the
//                walk count (uint32, clamped to MAX_PATHS when read, Gap2Seq.cpp:1058-1060), its node entry and
//                FOUR PREDECESSOR LINKS (slot nt = the predecessor that prepends base nt, the order of
//                Graph::predecessors, Graph.cpp:1316-1344).  The links are recorded while expanding forward,
//                so the backward subgraph pass (Gap2Seq.cpp:1173-1312) and the traceback (Gap2Seq.cpp:1437-1517)
//                never query the graph or the table again - they chase indices.
//   thit         for each right-flank target j and level d: the cell of (rnode(j), d), used by the found
//                check (Gap2Seq.cpp:1107-1159) and as the seeds of the backward pass.
//
// The group ends a gap by appending its marked cells (the union of all accepted walks and the traceback
// trellis, with safe bits when the subgraph is a plain level DAG) to a global output arena; the host then
// replays libc rand() in input order over those records (g2s_host.cpp).
#pragma once
#include "g2s_common.cuh"

namespace g2s {

// ------------------------------------------------------------------------------------------------ portability
#ifndef __CUDACC__
struct uint4 { uint32_t x, y, z, w; };
#endif
// shared-memory words written by other threads of the group are re-read through a volatile load
template <class T> G2S_D T ld_sh(const T *p) { return *(const volatile T *)p; }
#ifdef G2S_DEVICE_BUILD
template <class T> G2S_D T ld_cg(const T *p) { return __ldcg(p); }
G2S_D uint32_t a_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
// atomicAdd(p, 1) for counters that a whole CTA hammers (cell / node / border cursors in shared memory): the lanes
// that arrive together send ONE atomic and share out the range.  Same results as per-lane atomics, in another order.
G2S_D uint32_t a_inc(uint32_t *p) {
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(p, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
G2S_D unsigned long long a_add64(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
G2S_D uint32_t a_max(uint32_t *p, uint32_t v) { return atomicMax(p, v); }
G2S_D uint32_t a_or(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
G2S_D uint32_t a_cas(uint32_t *p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
G2S_D int32_t a_maxi(int32_t *p, int32_t v) { return atomicMax(p, v); }
G2S_D int32_t a_mini(int32_t *p, int32_t v) { return atomicMin(p, v); }
G2S_D uint64_t a_cas64(uint64_t *p, uint64_t cmp, uint64_t v) {
  return atomicCAS((unsigned long long *)p, (unsigned long long)cmp, (unsigned long long)v);
}
#else
template <class T> inline T ld_cg(const T *p) { return *p; }
inline uint32_t a_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
inline uint32_t a_inc(uint32_t *p) { return (*p)++; }
inline unsigned long long a_add64(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline uint32_t a_max(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
inline uint32_t a_or(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
inline uint32_t a_cas(uint32_t *p, uint32_t c, uint32_t v) { uint32_t o = *p; if (o == c) *p = v; return o; }
inline int32_t a_maxi(int32_t *p, int32_t v) { int32_t o = *p; if (v > o) *p = v; return o; }
inline int32_t a_mini(int32_t *p, int32_t v) { int32_t o = *p; if (v < o) *p = v; return o; }
inline uint64_t a_cas64(uint64_t *p, uint64_t c, uint64_t v) { uint64_t o = *p; if (o == c) *p = v; return o; }
#endif

// ------------------------------------------------------------------------------------------------ data layout
struct __attribute__((aligned(16))) NodeEnt {
  uint64_t key;        // canonical k-mer, KEY_EMPTY when free
  uint32_t rstamp;     // (level+1)*2 + strand of the node's last insertion into the right border
  uint32_t flags;      // NF_* bits; bits 8..31 = level+1 of the node's marked (Phase B) cell, 0 = none
  uint32_t fstamp[2];  // level+1 of the strand's most recent forward cell
  uint32_t fcell[2];   // that cell
};
enum { NF_IN_RS = 1, NF_HAS_F = 2, NF_TARGET = 4 };

enum {  // cell flag word
  CF_B = 1,        // on an accepted walk (vertex of the Phase B subgraph)
  CF_T = 2,        // on the traceback trellis
  CF_SEED = 4,     // has the edge to the sink
  CF_OMASK_SHIFT = 8
};

enum {  // PackRec.meta
  PM_B = 1, PM_T = 2, PM_SEED = 4, PM_START = 8, PM_SAFE = 16, PM_BASE_SHIFT = 8
};

struct PackRec {       // 32 B, one marked cell
  uint32_t pred[4];    // pack indices of the predecessors in A,C,T,G order of the prepended base (NONE32 = absent)
  uint32_t node;       // per-gap node id (index into the node table) - vertex identity of the subgraph
  uint32_t level;
  uint32_t meta;       // PM_* | last base << PM_BASE_SHIFT
  uint32_t count;      // clamped walk count of the cell
};

struct GapDev {        // 32 B input descriptor
  uint32_t flank_off;  // left flank (k+lmf chars) at flanks[flank_off], right flank (k+rmf chars) directly after it
  int32_t gap_len, gap_err, lmf, rmf;
  uint32_t flags;      // G2S_GAP_* of include/g2s.h
  uint32_t id;         // position in the caller's batch
  uint32_t rstart_off; // k+rmf chars whose k-mers START the right pass: fill_gap takes them from the END of kmer_right
                       // (Gap2Seq.cpp:871-885, 953-957) but the targets from its beginning (Gap2Seq.cpp:1119); the two
                       // differ only when kmer_right is longer than k+right_max_fuz (single-gap mode, Gap2Seq.cpp:239-252)
};

enum {  // GapOut.status (device side)
  GS_ORDER_DEPENDENT = 1, GS_OVERFLOW = 2, GS_GENERAL = 4, GS_OUT_FULL = 8, GS_INTERNAL = 16, GS_TRACED = 32, GS_DONE = 0x100
};

struct GapOut {        // 192 B
  uint32_t status;
  uint32_t count_fwd, count_all;
  uint32_t n_plen;
  int32_t plen[2];
  int32_t reached_fuz;
  uint32_t cells_right, cells_left, nodes;
  uint32_t npack;
  uint32_t start_pack[2];
  uint32_t nB, n_edges, n_src, n_sink;
  int32_t max_src_level, min_sink_level;
  uint32_t top_level;
  unsigned long long pack_off;
  uint32_t levels_r, levels_f;
  // GS_TRACED: the traceback had no choice to make, the device wrote the fill (case = safe / unsafe) itself
  unsigned long long fill_off;   // into the pass's fill arena
  uint32_t fill_len;             // characters, ends with the right k-mer (what the reference leaves at &fill[lmf - left_fuz])
  int32_t left_fuz;
  uint32_t rand_calls, pad0;
  uint32_t cyc[6];     // SM cycles (>> 4) spent in: init, right pass, forward pass, backward marking, pack, total
  uint32_t prof[8];    // forward pass detail (>> 4): scout, node phase, layout, found scan, general levels; [5] narrow batches,
                       // [6] narrow levels, [7] general levels
  uint32_t table_lookups;   // lookups in the global k-mer table (hash-gathers: one per scout segment / expanded successor / flank k-mer)
  uint32_t node_ops;        // find / find-or-insert operations on the gap's own node table
};

struct SliceLayout {   // byte offsets inside one scratch slice; identical for every slice of a launch
  unsigned long long bytes;
  uint32_t cap_cells, hash_log2, max_levels, max_fuz1;
  unsigned long long off_hash, off_ck, off_cc, off_cn, off_cp, off_cflags, off_packidx, off_packlist, off_packlevel,
      off_lstart, off_nB, off_nT, off_tcell, off_thit, off_lk, off_rk, off_rbk[2], off_rbm[2];
};

struct FillParams {
  Table table;
  const Table *tables;         // null, or one table per graph (per-gap filtered-reads mode): gap i uses tables[graph_of[i]]
  const uint32_t *graph_of;
  const char *flanks;
  const GapDev *gaps;
  const uint32_t *order;       // gap indices, longest first
  uint32_t n_gaps;
  GapOut *outs;
  unsigned char *scratch;
  SliceLayout lay;
  PackRec *pack;
  unsigned long long pack_cap;
  unsigned long long *pack_cursor;
  uint32_t *queue_head;
  char *fills;                 // fill arena of the device-side tracebacks (may be null: every gap is packed)
  unsigned long long fill_cap;
  unsigned long long *fill_cursor;
};

#ifndef G2S_CHAIN_LEN
#define G2S_CHAIN_LEN 128
#endif
static const int CHAIN_LEN = G2S_CHAIN_LEN;   // levels a narrow batch looks ahead
static const int NARROW_W = 4;     // widest border handled in narrow mode

static const int SCOUT_SEGS = 32;   // segments a chain may record per batch (a batch ends early beyond that)
struct ScoutSeg {      // one gather of a scout: tip x, then the hint's states on levels l0 .. l0 + h.len - 1
  uint64_t x;
  Hint h;              // the hinted bases, first base highest
  uint16_t l0;
  uint8_t end_mask, pad;   // end_mask: out-mask (walk orientation) of the state the hint ends at
};

struct NarrowBuf {     // look-ahead state of one pass (the right pass and the forward pass own one each, so that two warps
                       // can run them concurrently)
  uint32_t nnext;
  uint32_t rnext[3];   // right pass, general levels: entries of the next border, one counter per level in rotation (bit 31: the
                       // level could not insert a node), so that a level needs one barrier - nobody resets the counter being read
  uint32_t ch_len, ch_valid, ch_base, ch_cnt, nw_stop, nw_alive;
  uint64_t nw_x[NARROW_W], nw_x0[NARROW_W];
  uint32_t nw_m[NARROW_W], nw_cnt0[NARROW_W];
  Hint nw_h[NARROW_W];   // what is left of the tip's skip hint (generic scout)
  int32_t nw_fail[NARROW_W];
  int32_t nw_bad;              // first level of the batch that the narrow path cannot carry (start injection)
  int32_t nw_born[NARROW_W];   // first level (of the batch) on which the chain exists: 0, or where it split off its parent
  uint32_t nw_w;               // chains after the scout (splits add chains)
  uint32_t nw_par[NARROW_W];   // chain a split chain descends from
  uint32_t nw_n[CHAIN_LEN + 1], nw_off[CHAIN_LEN + 1], nw_run[CHAIN_LEN], nw_poff[CHAIN_LEN];
  uint32_t bk_p[NARROW_W][4];   // backward batches: predecessor cells of the batch's lowest level
  int32_t bk_len, scout_len;
  uint32_t nseg[NARROW_W];
  ScoutSeg seg[NARROW_W][SCOUT_SEGS];
  uint64_t ch_s[CHAIN_LEN][NARROW_W];
  uint32_t ch_a[CHAIN_LEN][NARROW_W], ch_e[CHAIN_LEN][NARROW_W];
};

struct GShared {       // per group, in shared memory
  uint32_t ncells, nnodes, status;
  uint32_t npack, nB, n_edges, n_src, n_sink;
  int32_t max_src_level, min_sink_level, max_seed_level;
  uint32_t cells_right;
  uint32_t n_plen;
  int32_t plen[2];
  int32_t reached_fuz;
  uint32_t count_fwd;
  int32_t found_level;
  unsigned long long count_all;
  unsigned long long pack_off;
  uint32_t fc1[MAX_FUZ + 1], fc2[MAX_FUZ + 1];
  uint32_t queue_item;
  unsigned long long prof[8];
  uint32_t table_lookups, node_ops;   // counters of the hash work (GapOut)
  int32_t tlo;          // first level that has thit entries: max(0, g + lmf - e)
  uint32_t trow;        // entries per target row: Dmax - tlo + 1
  // handed from the right pass / the forward pass to the rest of the gap (the two may run on different warps)
  int32_t levels_r, levels_f, top_level, fwd_overflow;
  // device-side traceback
  int32_t t_lo, t_left_fuz; uint32_t t_cnt, t_bad, t_ok;
  unsigned long long fill_off;
  unsigned long long cyc_right;
  NarrowBuf nbuf[2];    // [0] right pass, [1] forward pass and backward marking
};

struct Slice {         // typed view of one scratch slice
  NodeEnt *ent;
  uint32_t hmask, hcap;
  uint64_t *ck;
  uint32_t *cc, *cn;
  uint32_t *cp;        // 4 per cell
  uint32_t *cflags, *packidx, *packlist, *packlevel, *lstart, *nB, *thit;
  uint32_t *nT, *tcell;   // trellis (CF_T) cells per level and one of them: a level-by-level chain needs no random choice
  uint64_t *lk, *rk;   // oriented flank k-mers; bit 63 set = present in the graph; bits 62..: unused
  uint32_t cap_cells, max_levels;
  // right-border ping-pong buffers (oriented k-mer, in-mask); their own arrays: the right pass may run concurrently
  // with the forward pass
  uint64_t *rbk[2];
  unsigned char *rbm[2];
};

G2S_HD Slice make_slice(unsigned char *base, const SliceLayout &L) {
  Slice s;
  s.ent = (NodeEnt *)(base + L.off_hash);
  s.hcap = 1u << L.hash_log2;
  s.hmask = s.hcap - 1;
  s.ck = (uint64_t *)(base + L.off_ck);
  s.cc = (uint32_t *)(base + L.off_cc);
  s.cn = (uint32_t *)(base + L.off_cn);
  s.cp = (uint32_t *)(base + L.off_cp);
  s.cflags = (uint32_t *)(base + L.off_cflags);
  s.packidx = (uint32_t *)(base + L.off_packidx);
  s.packlist = (uint32_t *)(base + L.off_packlist);
  s.packlevel = (uint32_t *)(base + L.off_packlevel);
  s.lstart = (uint32_t *)(base + L.off_lstart);
  s.nB = (uint32_t *)(base + L.off_nB);
  s.nT = (uint32_t *)(base + L.off_nT);
  s.tcell = (uint32_t *)(base + L.off_tcell);
  s.thit = (uint32_t *)(base + L.off_thit);
  s.lk = (uint64_t *)(base + L.off_lk);
  s.rk = (uint64_t *)(base + L.off_rk);
  s.cap_cells = L.cap_cells;
  s.max_levels = L.max_levels;
  s.rbk[0] = (uint64_t *)(base + L.off_rbk[0]);
  s.rbk[1] = (uint64_t *)(base + L.off_rbk[1]);
  s.rbm[0] = base + L.off_rbm[0];
  s.rbm[1] = base + L.off_rbm[1];
  return s;
}

// ---- host-side geometry of a scratch slice (used by the engine when planning a pass)
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline SliceLayout make_layout(uint64_t cap_cells, uint32_t hash_log2, uint32_t max_levels, uint32_t max_fuz1) {
  SliceLayout L{};
  L.cap_cells = (uint32_t)cap_cells; L.hash_log2 = hash_log2; L.max_levels = max_levels; L.max_fuz1 = max_fuz1;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return (unsigned long long)r; };
  L.off_hash = take(((size_t)1 << hash_log2) * sizeof(NodeEnt));
  L.off_ck = take(cap_cells * 8);
  L.off_cc = take(cap_cells * 4);
  L.off_cn = take(cap_cells * 4);
  L.off_cp = take(cap_cells * 16);
  L.off_cflags = take(cap_cells * 4);
  L.off_packidx = take(cap_cells * 4);
  L.off_packlist = take(cap_cells * 4);
  L.off_packlevel = take(cap_cells * 4);
  L.off_lstart = take(((size_t)max_levels + 3) * 4);
  L.off_nB = take(((size_t)max_levels + 3) * 4);
  L.off_nT = take(((size_t)max_levels + 3) * 4);
  L.off_tcell = take(((size_t)max_levels + 3) * 4);
  L.off_thit = take((size_t)max_fuz1 * ((size_t)max_levels + 1) * 4);
  L.off_lk = take((size_t)max_fuz1 * 8);
  L.off_rk = take((size_t)max_fuz1 * 8);
  for (int q = 0; q < 2; q++) { L.off_rbk[q] = take(cap_cells * 8); L.off_rbm[q] = take(cap_cells); }
  L.bytes = o;
  return L;
}

inline uint32_t ceil_log2(uint64_t x) { uint32_t l = 0; while (((uint64_t)1 << l) < x) l++; return l; }


static const uint64_t FLANK_PRESENT = 1ULL << 63;
static const uint64_t FLANK_KMER = (1ULL << 62) - 1;

// ------------------------------------------------------------------------------------------------ node table
#if defined(G2S_HOST_EMUL) && defined(G2S_EMUL_STATS)
struct EmulStats { unsigned long long finds, find_probes, inserts, insert_probes, max_probe, width_hist[8], back_general_levels, back_chain_levels; };
inline EmulStats &emul_stats() { static EmulStats s; return s; }
#define G2S_STAT(x) do { x; } while (0)
#else
#define G2S_STAT(x) do { } while (0)
#endif

G2S_D uint32_t node_hash(uint64_t key, uint32_t hmask) { return (uint32_t)((key * 0x9E3779B97F4A7C15ULL) >> 40) & hmask; }

// The first 16 bytes of an entry (key, rstamp, flags) travel in one request: a lookup returns the node's flag word
// (right-set membership, target) without a second dependent round trip.
G2S_D void load_ent_head(const NodeEnt *p, uint64_t *key, uint32_t *flags) {
#ifdef G2S_DEVICE_BUILD
  const uint4 v = __ldcg((const uint4 *)p);
  *key = (uint64_t)v.x | ((uint64_t)v.y << 32);
  *flags = v.w;
#else
  *key = p->key; *flags = p->flags;
#endif
}

G2S_D uint32_t node_find(const Slice &S, uint64_t key, uint32_t *flags) {
  uint32_t h = node_hash(key, S.hmask);
  G2S_STAT(emul_stats().finds++);
  for (;;) {
    G2S_STAT(emul_stats().find_probes++);
    uint64_t k0;
    load_ent_head(&S.ent[h], &k0, flags);
    if (k0 == key) return h;
    if (k0 == KEY_EMPTY) return NONE32;
    h = (h + 1) & S.hmask;
  }
}

// Find or insert.  Most inserted keys are new (a walk rarely revisits a node), so the probe IS the compare-and-swap:
// one round trip for a new node; an existing one costs a second (its flag word).  *flags receives the flag word as
// of the lookup (0 for a new node).
G2S_D uint32_t node_insert(const Slice &S, GShared &sh, uint64_t key, uint32_t *flags) {
  uint32_t h = node_hash(key, S.hmask);
  G2S_STAT(emul_stats().inserts++);
  if (ld_sh(&sh.nnodes) >= S.hcap - (S.hcap >> 2)) { a_or(&sh.status, GS_OVERFLOW); return NONE32; }
  for (;;) {
    G2S_STAT(emul_stats().insert_probes++);
    const uint64_t old = a_cas64(&S.ent[h].key, KEY_EMPTY, key);
    if (old == KEY_EMPTY) { a_inc(&sh.nnodes); *flags = 0; return h; }
    if (old == key) { *flags = ld_cg(&S.ent[h].flags); return h; }
    h = (h + 1) & S.hmask;
  }
}
G2S_D uint32_t node_insert(const Slice &S, GShared &sh, uint64_t key) { uint32_t f; return node_insert(S, sh, key, &f); }

// Sum of v over the calling warp, added to *p by one lane (all 32 lanes must call; device only - the host emulation
// has groups of one thread).  For per-level totals that every marked cell of a wide level would otherwise add itself.
G2S_D void a_add_warp(uint32_t *p, uint32_t v) {
#ifdef __CUDA_ARCH__
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(p, v);
#else
  if (v) *p += v;
#endif
}

// Has the gap outgrown its scratch slice?  The answer steers collective control flow (loops with group barriers), so
// it must be the same for every thread of the group - but in the thin tier the OTHER warp of the pair may set the bit
// at any moment, and lanes reading the word one after the other could disagree.  G::any() makes it a group decision.
template <class G>
G2S_D bool group_status_any(GShared &sh, uint32_t bits) { return G::any_word(&sh.status, bits); }
template <class G>
G2S_D bool group_overflow(GShared &sh) { return group_status_any<G>(sh, GS_OVERFLOW); }

G2S_D unsigned long long now_cycles() {
#ifdef __CUDA_ARCH__
  return (unsigned long long)clock64();
#else
  return 0;
#endif
}

G2S_D uint32_t clamp_paths(uint32_t c) { return c > MAX_PATHS ? MAX_PATHS : c; }

// create the forward cell (state s of node e) on level c; the caller owns the (e,strand,level) stamp
template <class G>
G2S_D uint32_t new_cell(const Slice &S, GShared &sh, const GapDev &gd, uint64_t s, uint32_t e, uint32_t strand,
                        uint32_t adj, uint32_t c, uint32_t count, uint32_t nflags) {
  uint32_t idx = a_inc(&sh.ncells);
  if (idx >= S.cap_cells) { a_or(&sh.status, GS_OVERFLOW); return NONE32; }
  S.ck[idx] = s;
  S.cc[idx] = count;
  S.cn[idx] = e;
  S.cp[4 * idx + 0] = NONE32; S.cp[4 * idx + 1] = NONE32; S.cp[4 * idx + 2] = NONE32; S.cp[4 * idx + 3] = NONE32;
  S.cflags[idx] = out_mask(adj, strand) << CF_OMASK_SHIFT;
  S.ent[e].fcell[strand] = idx;
  if ((nflags & NF_TARGET) && (int)c >= sh.tlo) {
    for (int j = 0; j <= gd.rmf; j++)
      if ((S.rk[j] & FLANK_KMER) == s) S.thit[(size_t)j * sh.trow + (c - (uint32_t)sh.tlo)] = idx + 1;
  }
  return idx;
}

// same, into a cell the caller has already reserved (warp-aggregated reservation in the pair step below), with the
// creating pair's walk count and predecessor link
G2S_D void new_cell_at(const Slice &S, const GShared &sh, const GapDev &gd, uint32_t idx, uint64_t s, uint32_t e, uint32_t strand,
                       uint32_t adj, uint32_t c, uint32_t count, uint32_t nflags, uint32_t u0, uint32_t pred) {
  S.ck[idx] = s;
  S.cc[idx] = count;
  S.cn[idx] = e;
  // predecessor links by prepended base: the creating pair's own link is there from the start
  const uint4 links = {u0 == 0 ? pred : NONE32, u0 == 1 ? pred : NONE32, u0 == 2 ? pred : NONE32, u0 == 3 ? pred : NONE32};
  *(uint4 *)&S.cp[4 * idx] = links;
  S.cflags[idx] = out_mask(adj, strand) << CF_OMASK_SHIFT;
  S.ent[e].fcell[strand] = idx;
  if ((nflags & NF_TARGET) && (int)c >= sh.tlo) {
    for (int j = 0; j <= gd.rmf; j++)
      if ((S.rk[j] & FLANK_KMER) == s) S.thit[(size_t)j * sh.trow + (c - (uint32_t)sh.tlo)] = idx + 1;
  }
}

#ifdef G2S_DEVICE_BUILD
// ---- the pair step of a wide level.  A border element has up to four neighbours and each of them costs a chain of
// dependent memory operations (k-mer table gather, node-table operation, stamp, cell).  With one thread per element a
// warp walks that chain once per base - four times in a row, because some lane always has the base - and the level
// step of a wide gap is nothing but this latency (profiles/r2_summary.md).  Here a warp takes C <= 32 elements, lays
// their (element, base) pairs out over its lanes with a prefix sum of the neighbour counts and walks the chain once
// per 32 PAIRS: one pass for most warps.
struct PairSrc {
  uint32_t incl, excl, total;
};
G2S_D PairSrc pair_scan(uint32_t cnt) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t v = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (uint32_t)d) v += t;
  }
  PairSrc r;
  r.incl = v; r.excl = v - cnt; r.total = __shfl_sync(0xffffffffu, v, 31);
  return r;
}
// lane that owns pair p (all lanes call): the last lane whose exclusive prefix is <= p - lanes without pairs share the
// prefix of the next lane, so the last one of a run is the one that has pairs.  Five shuffles.
G2S_D uint32_t pair_owner(const PairSrc &ps, uint32_t p, uint32_t C) {
  (void)C;
  uint32_t src = 0;
#pragma unroll
  for (uint32_t st = 16; st; st >>= 1) {
    const uint32_t t = __shfl_sync(0xffffffffu, ps.excl, src + st);
    if (t <= p) src += st;
  }
  return src;
}
G2S_D uint32_t nth_bit(uint32_t m, uint32_t r) {   // index of the r-th (0-based) set bit of a 4-bit mask
  for (; r; r--) m &= m - 1;
  return (uint32_t)__ffs((int)m) - 1u;
}
G2S_D uint64_t shfl64(uint64_t v, uint32_t src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
// 16-byte compare-and-swap on the head of a node entry (key, rstamp, flags): claims a free entry for `key` or returns
// what is there - the flag word included - in ONE round trip (ATOMG.E.CAS.128)
G2S_D void ent_cas_head(NodeEnt *p, uint64_t key, uint64_t *okey, uint32_t *oflags) {
  uint64_t olo, ohi;
  asm volatile("{\n\t.reg .b128 c, s, d;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 s, {%5, %6};\n\tatom.global.cas.b128 d, [%2], c, s;\n\tmov.b128 {%0, %1}, d;\n\t}"
               : "=l"(olo), "=l"(ohi) : "l"(p), "l"(KEY_EMPTY), "l"(0ULL), "l"(key), "l"(0ULL) : "memory");
  *okey = olo; *oflags = (uint32_t)(ohi >> 32);
}
// The two first probes of a neighbour - its slot of the global k-mer table and its entry of the gap's node table -
// depend on nothing but the canonical k-mer and are issued together.  INSERT: find or insert; else find only.
// Returns the node entry (NONE32: absent / table full) and the adjacency byte; *nflags = the node's flag word.
template <bool INSERT>
G2S_D uint32_t table_and_node(const Table &T, const Slice &S, GShared &sh, uint64_t key, uint32_t nnodes_seen, uint32_t *adj, uint32_t *nflags) {
  uint64_t ht = mix64(key) & T.cap_mask;
  uint32_t hn = node_hash(key, S.hmask);
  Slot ts = load_slot(T.slots + ht);
  uint64_t nk; uint32_t nf;
  const bool full = INSERT && nnodes_seen >= S.hcap - (S.hcap >> 2);
  if (INSERT && !full) ent_cas_head(&S.ent[hn], key, &nk, &nf);
  else load_ent_head(&S.ent[hn], &nk, &nf);
  while (ts.key != key && ts.key != KEY_EMPTY) { ht = (ht + 1) & T.cap_mask; ts = load_slot(T.slots + ht); }
  if (ts.key == KEY_EMPTY) { a_or(&sh.status, GS_INTERNAL); return NONE32; }   // the adjacency byte said it exists
  *adj = (uint32_t)(ts.val & 0xFF);
  for (;;) {
    if (nk == key) break;
    if (nk == KEY_EMPTY) {
      if (!INSERT) return NONE32;
      if (full) { a_or(&sh.status, GS_OVERFLOW); return NONE32; }
      a_inc(&sh.nnodes); nf = 0; break;       // the compare-and-swap above claimed the entry
    }
    hn = (hn + 1) & S.hmask;
    if (INSERT && !full) ent_cas_head(&S.ent[hn], key, &nk, &nf);
    else load_ent_head(&S.ent[hn], &nk, &nf);
  }
  *nflags = nf;
  return hn;
}
// reserve one item per lane that asks for it with ONE atomic per warp (all 32 lanes call)
G2S_D uint32_t warp_reserve(uint32_t *counter, bool want) {
  const uint32_t m = __ballot_sync(0xffffffffu, want);
  if (!m) return 0;
  const uint32_t lane = threadIdx.x & 31, leader = (uint32_t)__ffs((int)m) - 1u;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
#endif

// thit holds, for target j and level d >= tlo, the cell of (rnode(j), d) + 1.  Levels below tlo = g + lmf - e are
// never looked up: the found check reads lengths >= g + lmf - err, the seeds levels >= lmf + g - e.
G2S_D uint32_t thit_get(const Slice &S, const GShared &sh, int j, int level) {
  if (level < sh.tlo || (uint32_t)(level - sh.tlo) >= sh.trow) return 0;
  return ld_cg(&S.thit[(size_t)j * sh.trow + (uint32_t)(level - sh.tlo)]);
}
G2S_D uint32_t thit_count(const Slice &S, const GShared &sh, int j, int level) {
  const uint32_t t = thit_get(S, sh, j, level);
  return t ? clamp_paths(ld_cg(&S.cc[t - 1])) : 0;
}

// The first-found check of level c (Gap2Seq.cpp:1108-1154), evaluated by ONE thread.  Returns true when found.
G2S_D bool found_at_level(const Slice &S, const GShared &sh, const GapDev &gd, int c, int *J, uint32_t *c1, uint32_t *c2, int *l1, int *l2) {
  const int err = c - gd.gap_len - gd.lmf - gd.rmf;
  for (int j = 0; j <= gd.rmf; j++) {
    int len1 = gd.gap_len + gd.lmf + j + err, len2 = gd.gap_len + gd.lmf + j - err;
    uint32_t a = thit_count(S, sh, j, len1);
    uint32_t b = (len2 != len1 && len2 >= 0) ? thit_count(S, sh, j, len2) : 0;
    if (a | b) { *J = j; *c1 = a; *c2 = b; *l1 = len1; *l2 = len2; return true; }
  }
  return false;
}

// Collective: first level in [lo, hi] whose found check succeeds (one level per thread); fills the shared found
// state like the per-level check does.  Returns true when a level was found.
template <class G>
G2S_D bool found_scan(const Slice &S, const GapDev &gd, GShared &sh, int lo, int hi) {
  for (int c = lo + G::rank(); c <= hi; c += G::size()) {
    int J = 0, l1 = 0, l2 = 0; uint32_t c1 = 0, c2 = 0;
    if (found_at_level(S, sh, gd, c, &J, &c1, &c2, &l1, &l2)) { a_mini(&sh.found_level, c); break; }
  }
  G::sync();
  const int fl = ld_sh(&sh.found_level);
  const bool found = fl <= hi;
  if (found && G::rank() == 0) {
    int J = 0, l1 = 0, l2 = 0; uint32_t c1 = 0, c2 = 0;
    found_at_level(S, sh, gd, fl, &J, &c1, &c2, &l1, &l2);
    sh.reached_fuz = J;
    uint64_t tot = 0; uint32_t n = 0;
    if (c1) { tot += c1; sh.plen[n++] = l1; }
    if (c2) { tot += c2; sh.plen[n++] = l2; }
    sh.n_plen = n;
    sh.count_fwd = tot > MAX_PATHS ? MAX_PATHS : (uint32_t)tot;
  }
  G::sync();
  return found;
}

// ---- narrow mode.  While a border holds at most NARROW_W states and each of them has at most one neighbour, the
// only true dependency between levels is the adjacency byte of the next k-mer of each chain (it decides which base
// comes after it).  One scout thread per border state therefore walks up to CHAIN_LEN levels ahead through the
// global table - one dependent gather per level, the chains' gathers in flight together - and the group then does
// everything else for those levels in bulk: node-table work and prune tests in parallel (one thread per level and
// chain), a short serial pass over shared memory that merges chains meeting in one state (bubbles closing), adds
// their counts and lays out the cells, and a parallel commit of stamps / target hits.  Branching levels (a state
// with two neighbours) and the levels that inject flank starts stay on the level-by-level path.  Results are
// identical to it.

// Scout phase shared by both passes.  Returns the number of levels looked ahead (0: a border state branches).
// ch_s[l][t] / ch_a[l][t] receive chain t's state and (adj | strand << 8 | 1 << 16) on level l, ch_a == 0: no state.
#ifdef G2S_DEVICE_BUILD
// Warp flavour of the scout.  Lane t owns chain t and runs nothing but the chain of dependent gathers: a gather
// returns the out-mask of the tip and its skip hint, i.e. the next j <= HINT_MAX states without looking at the table;
// the lane records that as ONE segment (tip, hint, first level), jumps to the node the hint ends at and gathers
// again.  Chains advance independently of each other.  The states of all segments are then computed by every thread
// of the group in parallel (expand_segments): the per-level work is off the serial path.
// SPLIT: a tip with several neighbours does not end the look-ahead while free chain slots remain: the extra
// neighbours become new chains, born on that level, that carry their parent's initial count (forward pass; the layout
// finds their predecessor cell through nw_par).  Without a free slot the batch ends before the level of the branch.
template <bool FORWARD, bool SPLIT>
__device__ __forceinline__ int narrow_scout_warp(const Table &T, GShared &sh, NarrowBuf &nb, int w, int maxlen) {
  const int k = T.k;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  uint64_t x = 0; uint32_t mask = 0;
  int lv = 0;                 // levels of the batch this chain has covered: its next state belongs to level lv
  int stop = maxlen;          // the batch must end before this level because of this chain
  int endlv = -1;             // first level without a state of this chain (dead end), -1: none
  int nseg = 0;
  bool active = lane < w;
  if (active) { x = nb.nw_x[lane]; mask = nb.nw_m[lane]; }
  Hint hint{0, 0, 0};         // hint of the tip (len 0: none)
  while (__any_sync(FULL, active)) {
    // ---- branches: split into new chains while slots remain
    const bool branch = active && (mask & (mask - 1)) != 0;
    const uint32_t bvote = __ballot_sync(FULL, branch);
    if (bvote) {
      int slot = w;
      for (int p = 0; p < NARROW_W; p++) {                           // parents in chain order
        if (!((bvote >> p) & 1)) continue;
        const uint32_t mp = __shfl_sync(FULL, mask, p);
        const uint64_t xp = __shfl_sync(FULL, x, p);
        const int lp = __shfl_sync(FULL, lv, p);
        const int extra = __popc(mp) - 1;
        if (!SPLIT || slot + extra > NARROW_W) {                     // no room: the batch ends before this level
          if (lane == p) { stop = lv; active = false; }
          continue;
        }
        uint32_t rest = mp & (mp - 1);                               // every neighbour but the first
        if (lane == p) { mask = mp & (0u - mp); hint.len = 0; }      // the parent keeps the first
        while (rest) {
          const uint32_t bit = rest & (0u - rest);
          rest &= rest - 1;
          if (lane == slot) { x = xp; mask = bit; hint.len = 0; lv = lp; active = true; stop = maxlen; endlv = -1; nseg = 0; }
          if (lane == p) { nb.nw_x0[slot] = nb.nw_x0[p]; nb.nw_cnt0[slot] = nb.nw_cnt0[p]; nb.nw_born[slot] = lp; nb.nw_par[slot] = (uint32_t)p; }
          slot++;
        }
      }
      w = slot;
    }
    if (active) {
      if (mask == 0) { endlv = lv; active = false; }                 // dead end: no state from level lv on
      else if (lv >= maxlen) active = false;
      else if (nseg >= SCOUT_SEGS) { stop = lv; active = false; }    // out of segment slots: a shorter batch
    }
    if (active) {
      if (hint.len == 0) { hint.lo = mask == 1 ? 0 : mask == 2 ? 1 : mask == 4 ? 2 : 3; hint.hi = 0; hint.len = 1; }   // no hint: one known step
      // the node the hint ends at
      const uint64_t xe = hint_state(x, hint, hint.len, T.kmask);
      const Canon ce = canonical(xe, k);
      Slot sl;
      uint32_t emask = 0;
      const bool hit = table_find_slot(T, ce.key, &sl) >= 0;
      if (!hit) a_or(&sh.status, GS_INTERNAL);
      else emask = out_mask((uint32_t)(sl.val & 0xFF), ce.strand);
      ScoutSeg &sg = nb.seg[lane][nseg++];
      sg.x = x; sg.h = hint; sg.l0 = (uint16_t)lv; sg.end_mask = (uint8_t)emask;
      x = xe; lv += (int)hint.len; mask = emask;
      hint.len = 0;
      if (hit && mask != 0 && !(mask & (mask - 1))) hint = hint_of(sl, ce.strand);
    }
  }
  if (lane < NARROW_W) nb.nseg[lane] = lane < w ? (uint32_t)nseg : 0u;
  // length of the batch: up to the first level a chain has to stop before; if every chain ran into a dead end
  // earlier, one empty level is kept so that the caller sees the border die
  int st = lane < w ? stop : maxlen;
  int en = lane < w ? (endlv >= 0 ? endlv : 0x7fffffff) : -1;       // lanes without a chain do not keep the batch alive
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    st = min(st, __shfl_xor_sync(FULL, st, o));
    en = max(en, __shfl_xor_sync(FULL, en, o));
  }
  int len = st;
  if (en != 0x7fffffff && en < len) len = en + 1 < maxlen ? en + 1 : maxlen;
  if (lane == 0) { nb.nw_w = (uint32_t)w; nb.scout_len = len; }
  __syncwarp();
  return len;
}

// States of the scouted segments, one (level, chain) per thread: binary search for the segment that covers the level.
// ch_a of the covered levels was zeroed before.
template <class G, bool FORWARD>
__device__ __forceinline__ void expand_segments(const Table &T, NarrowBuf &nb, int w, int len) {
  const int k = T.k;
  for (int q = G::rank(); q < len * w; q += G::size()) {
    const int lvl = q / w, t = q - lvl * w;
    const int ns = (int)nb.nseg[t];
    if (ns == 0 || lvl < (int)nb.seg[t][0].l0) continue;            // not born yet (split chains start later)
    int lo = 0, hi = ns - 1;                                          // last segment with l0 <= lvl
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((int)nb.seg[t][mid].l0 <= lvl) lo = mid; else hi = mid - 1; }
    const ScoutSeg sg = nb.seg[t][lo];
    const int j = lvl - (int)sg.l0 + 1;                               // steps past the segment's tip
    if (j > (int)sg.h.len) continue;                                  // the chain ended before this level
    const uint64_t y = hint_state(sg.x, sg.h, (uint32_t)j, T.kmask);
    const Canon cy = canonical(y, k);
    const uint32_t nmask = j < (int)sg.h.len ? 1u << hint_base(sg.h, (uint32_t)j) : (uint32_t)sg.end_mask;
    if (FORWARD) { nb.ch_s[lvl][t] = y; nb.ch_a[lvl][t] = nmask | (cy.strand << 8) | (1u << 16); }
    else { nb.ch_s[lvl][t] = revcomp(y, k); nb.ch_a[lvl][t] = swap2(nmask) | ((cy.strand ^ 1u) << 8) | (1u << 16); }
  }
}
#endif

// hash-gather accounting: the warp scout makes one table lookup per recorded segment
G2S_D void count_scout_lookups(GShared &sh, const NarrowBuf &nb, int w) {
#ifdef G2S_DEVICE_BUILD
  uint32_t n = 0;
  for (int t = 0; t < w && t < NARROW_W; t++) n += nb.nseg[t];
  a_add(&sh.table_lookups, n);
#endif
}

template <class G, bool FORWARD>
G2S_D int narrow_scout(const Table &T, GShared &sh, NarrowBuf &nb, int w, int maxlen) {
#ifdef G2S_DEVICE_BUILD
  if (G::kWarp || G::kWarpScout) {   // the group's first warp scouts, then every thread expands the segments
    G::sync();
    if (G::rank() < 32) narrow_scout_warp<FORWARD, true>(T, sh, nb, w, maxlen);
    G::sync();
    const int len = ld_sh(&nb.scout_len), w2 = (int)ld_sh(&nb.nw_w);
    for (int q = G::rank(); q < len * NARROW_W; q += G::size()) nb.ch_a[q / NARROW_W][q % NARROW_W] = 0;
    G::sync();
    expand_segments<G, FORWARD>(T, nb, w2, len);
    G::sync();
    return len;
  }
#endif
  // Per-chain state lives in shared memory so that any group size, including the size-1 host emulation, can play the
  // w scouts: nw_x = the chain's tip in WALK orientation (forward pass: the state itself; right pass: its reverse
  // complement, so that walking predecessors is walking successors), nw_m = its out-mask in that orientation
  // (0xFF = chain ended), nw_h = what is left of the tip's skip hint.  A gather happens only where a hint
  // ends; between gathers a level costs a shift and an insert.
  const int k = T.k, rank = G::rank(), gsz = G::size();
  if (rank == 0) { nb.nw_stop = 0; nb.nw_alive = 0; nb.nw_w = (uint32_t)w; }
  G::sync();
  int len = 0;
  for (int l = 0; l < maxlen; l++) {
    for (int t = rank; t < w; t += gsz) {
      const uint32_t mask = nb.nw_m[t];
      if (mask == 0xFF) continue;
      if (mask == 0) nb.nw_m[t] = 0xFF;                              // dead end
      else if (mask & (mask - 1)) nb.nw_stop = 1;                    // two neighbours: not a chain any more
      else nb.nw_alive = (uint32_t)l + 1;
    }
    G::sync();
    if (ld_sh(&nb.nw_stop) || ld_sh(&nb.nw_alive) < (uint32_t)l + 1) {
      // a branch ends the look-ahead BEFORE this level; "all chains ended" keeps the (empty) level so that the
      // caller records the border as dead
      if (!ld_sh(&nb.nw_stop)) { for (int t = rank; t < w; t += gsz) nb.ch_a[l][t] = 0; len = l + 1; }
      break;
    }
    for (int t = rank; t < w; t += gsz) {
      const uint32_t mask = nb.nw_m[t];
      if (mask == 0xFF) { nb.ch_a[l][t] = 0; continue; }
      const uint32_t nt = mask == 1 ? 0 : mask == 2 ? 1 : mask == 4 ? 2 : 3;
      const uint64_t y = succ_kmer(nb.nw_x[t], nt, T);
      const Canon cy = canonical(y, k);
      uint32_t nmask;
      if (nb.nw_h[t].len > 1) {        // y is an interior node of the hint: exactly one successor, its base is known
        nb.nw_h[t].len--;               // (the remaining bases are the low 2 * len bits)
        nmask = 1u << ((uint32_t)hint_shr(nb.nw_h[t], 2 * (nb.nw_h[t].len - 1)) & 3u);
      } else {                         // the hint ends at y (or there is none yet): gather y
        Slot sl;
        if (table_find_slot(T, cy.key, &sl) < 0) { a_or(&sh.status, GS_INTERNAL); nb.nw_m[t] = 0xFF; nb.ch_a[l][t] = 0; continue; }
        nmask = out_mask((uint32_t)(sl.val & 0xFF), cy.strand);
        nb.nw_h[t] = hint_of(sl, cy.strand);
      }
      // emit the state of this level in the orientation of the pass, with its mask in the direction of the pass
      if (FORWARD) { nb.ch_s[l][t] = y; nb.ch_a[l][t] = nmask | (cy.strand << 8) | (1u << 16); }
      else { nb.ch_s[l][t] = revcomp(y, k); nb.ch_a[l][t] = swap2(nmask) | ((cy.strand ^ 1u) << 8) | (1u << 16); }
      nb.nw_x[t] = y; nb.nw_m[t] = nmask;
    }
    len = l + 1;
    G::sync();
  }
  G::sync();
  return len;
}

// Right pass: advances the border S.rbk[cur][0..n) (n <= NARROW_W) by up to CHAIN_LEN levels starting with level c.
// Returns the number of levels done; *n_out receives the size of the new border.
template <class G>
G2S_D int narrow_right(const Table &T, const Slice &S, GShared &sh, NarrowBuf &nb, int cur, int c, int right_half, int n, uint32_t *n_out) {
  const int k = T.k, rank = G::rank(), gsz = G::size();
  for (int t = rank; t < n; t += gsz) { nb.nw_x[t] = revcomp(S.rbk[cur][t], k); nb.nw_m[t] = swap2(S.rbm[cur][t]); nb.nw_h[t] = Hint{0, 0, 0}; }
  const int maxlen = right_half - c + 1 < CHAIN_LEN ? right_half - c + 1 : CHAIN_LEN;
  const int len = narrow_scout<G, false>(T, sh, nb, n, maxlen);
  if (len == 0) return 0;
  n = (int)ld_sh(&nb.nw_w);   // the scout splits chains at branches; a chain has no state before its birth level
  if (rank == 0) { nb.ch_cnt = 0; nb.nnext = 0; count_scout_lookups(sh, nb, n); }
  G::sync();
  uint32_t my_ops = 0;
  for (int q = rank; q < len * n; q += gsz) {
    const int l = q / n, t = q - l * n;
    if (!nb.ch_a[l][t]) continue;
    const Canon cp = canonical(nb.ch_s[l][t], k);
    const uint32_t en = node_insert(S, sh, cp.key);
    my_ops++;
    if (en == NONE32) continue;
    a_or(&S.ent[en].flags, NF_IN_RS);
    a_max(&S.ent[en].rstamp, (uint32_t)(c + l + 1) * 2 + cp.strand);
  }
  // distinct states per level = cells of the right table (chains that met are counted once); one level per thread
  for (int l = rank; l < len; l += gsz) {
    uint32_t ns = 0;
#pragma unroll
    for (int t = 0; t < NARROW_W; t++) {
      if (t >= n || !nb.ch_a[l][t]) continue;
      const uint64_t s = nb.ch_s[l][t];
      bool dup = false;
#pragma unroll
      for (int u = 0; u < NARROW_W; u++) {
        if (u >= t || !nb.ch_a[l][u]) continue;
        if (nb.ch_s[l][u] == s) dup = true;
        else if (nb.ch_s[l][u] == revcomp(s, k)) a_or(&sh.status, GS_ORDER_DEPENDENT);   // both strands on one level (Q3)
      }
      if (dup) continue;
      if (l == len - 1) {
        S.rbk[cur][ns] = s;
        S.rbm[cur][ns] = (unsigned char)(nb.ch_a[l][t] & 0xFF);   // already the in-mask of the emitted state
      }
      ns++;
    }
    a_add(&nb.ch_cnt, ns);
    if (l == len - 1) nb.nnext = ns;
  }
  if (my_ops) a_add(&sh.node_ops, my_ops);
  G::sync();
  if (rank == 0) sh.cells_right += nb.ch_cnt;
  *n_out = ld_sh(&nb.nnext);
  G::sync();
  return len;
}

// Forward pass: the border of level c-1 is cells [lo, lo+w), w <= NARROW_W.  Creates the cells of levels
// c .. c+ret-1 (the last returned level may be empty: the border died there).
//
// Chains never split in narrow mode, they only end (dead end, pruned) or meet (equal states, bubbles closing), and
// once met they stay equal.  So everything about level l follows from level l's candidates alone:
//   alive(t, l)  = l < first failing level of chain t
//   cell of t    = the lowest alive chain with the same state on level l
//   its count    = sum of the INITIAL counts of the alive chains with that state (clamping commutes)
//   its links    = the cells of those chains on level l-1
// which lets one thread per level lay out and write that level's cells once a prefix sum has placed the levels.
template <class G>
G2S_D int narrow_forward(const Table &T, const Slice &S, GShared &sh, NarrowBuf &nb, const GapDev &gd, uint32_t lo, int w, int c, int Dmax,
                         int prune_from) {
  const int lmf = gd.lmf;
  const int k = T.k, rank = G::rank(), gsz = G::size();
  for (int t = rank; t < NARROW_W; t += gsz) nb.nw_born[t] = 0;
  for (int t = rank; t < w; t += gsz) {
    const uint64_t x = ld_cg(&S.ck[lo + t]);
    nb.nw_x[t] = x; nb.nw_x0[t] = x;
    nb.nw_m[t] = (ld_cg(&S.cflags[lo + t]) >> CF_OMASK_SHIFT) & 15u;
    nb.nw_cnt0[t] = clamp_paths(ld_cg(&S.cc[lo + t]));
    nb.nw_h[t] = Hint{0, 0, 0};
  }
  const int maxlen = Dmax - c + 1 < CHAIN_LEN ? Dmax - c + 1 : CHAIN_LEN;
  const unsigned long long tp0 = now_cycles();
  const int len = narrow_scout<G, true>(T, sh, nb, w, maxlen);
  const unsigned long long tp1 = now_cycles();
  if (rank == 0) sh.prof[0] += tp1 - tp0;
  if (len == 0) return 0;
  w = (int)ld_sh(&nb.nw_w);   // the scout may have split chains
  for (int t = rank; t < w; t += gsz) nb.nw_fail[t] = len;
  if (rank == 0) { nb.nw_bad = len; count_scout_lookups(sh, nb, w); }
  G::sync();
  uint32_t my_ops = 0;
  // node table + prune test, one (level, chain) per thread: ch_e = node entry; first failing level per chain
  for (int q = rank; q < len * w; q += gsz) {
    const int l = q / w, t = q - l * w;
    uint32_t en = NONE32;
    if (l < nb.nw_born[t]) { nb.ch_e[l][t] = NONE32; continue; }   // not born yet: neither a state nor a failure
    if (nb.ch_a[l][t]) {
      const Canon cs = canonical(nb.ch_s[l][t], k);
      uint32_t nf = 0;
      if (c + l < prune_from) en = node_insert(S, sh, cs.key, &nf);
      else { en = node_find(S, cs.key, &nf); if (en != NONE32 && !(nf & NF_IN_RS)) en = NONE32; }
      my_ops++;
      if (nf & NF_TARGET) nb.ch_a[l][t] |= 1u << 17;
    }
    nb.ch_e[l][t] = en;
    if (en == NONE32) a_mini(&nb.nw_fail[t], l);
  }
  if (my_ops) a_add(&sh.node_ops, my_ops);
  G::sync();
  // the scout splits chains before it knows where chains get pruned: a chain whose parent failed before the split
  // level descends from a state that has no cell, so it never lives (parents have lower indices than their children)
  if (rank == 0)
    for (int t = 1; t < w; t++)
      if (nb.nw_born[t] > 0 && nb.nw_fail[nb.nw_par[t]] < nb.nw_born[t]) nb.nw_fail[t] = 0;
  G::sync();
  const unsigned long long tp2 = now_cycles();
  if (rank == 0) sh.prof[1] += tp2 - tp1;
  if (group_overflow<G>(sh)) return 0;
  // distinct cells per level
  for (int l = rank; l < len; l += gsz) {
    uint32_t nn = 0;
#pragma unroll
    for (int t = 0; t < NARROW_W; t++) {
      if (t >= w || l >= nb.nw_fail[t] || l < nb.nw_born[t]) continue;
      bool dup = false;
#pragma unroll
      for (int u = 0; u < NARROW_W; u++)
        if (u < t && l < nb.nw_fail[u] && l >= nb.nw_born[u] && nb.ch_s[l][u] == nb.ch_s[l][t]) dup = true;
      if (!dup) nn++;
    }
    nb.nw_n[l] = nn;
    // Levels that inject a staggered left start (Gap2Seq.cpp:1082-1105: the start's count is OVERWRITTEN with 1): the
    // batch can carry such a level only if the overwrite changes nothing - the start k-mer is the state of living
    // chains whose counts add up to 1, i.e. of exactly one chain that counts one walk (the usual case: the flank
    // is a path of the graph).  Otherwise the batch ends before that level and the level-by-level path injects.
    if (c + l <= lmf) {
      const uint64_t sk = S.lk[c + l];
      if (sk & FLANK_PRESENT) {
        uint32_t sum = 0;
#pragma unroll
        for (int t = 0; t < NARROW_W; t++)
          if (t < w && l < nb.nw_fail[t] && l >= nb.nw_born[t] && nb.ch_s[l][t] == (sk & FLANK_KMER)) sum += nb.nw_cnt0[t] > 1 ? 2u : nb.nw_cnt0[t];
        if (sum != 1) a_mini(&nb.nw_bad, l);
      }
    }
  }
  G::sync();
#ifdef G2S_DEVICE_BUILD
  if (rank < 32) {   // place the levels (first warp): where the batch ends, prefix sum of the level widths
    const unsigned FULL = 0xffffffffu;
    const int LPL = (CHAIN_LEN + 31) / 32;
    const int bad = ld_sh(&nb.nw_bad);
    uint32_t nn[LPL];
    int my_end = 0x7fffffff;   // number of levels laid out if the batch ends inside this lane's levels
#pragma unroll
    for (int j = 0; j < LPL; j++) {
      const int l = rank * LPL + j;
      nn[j] = l < len ? nb.nw_n[l] : 0u;
      if (l < len && my_end == 0x7fffffff) {
        if (l >= bad) my_end = l;                  // a level the batch cannot carry: it ends before it
        else if (nn[j] == 0) my_end = l + 1;       // the border died: the empty level is kept
      }
    }
    int done = my_end;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) done = min(done, __shfl_xor_sync(FULL, done, o));
    if (done > len) done = len;
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < LPL; j++) { if (rank * LPL + j >= done) nn[j] = 0; mine += nn[j]; }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (rank >= o) incl += v; }
    uint32_t run = incl - mine;
#pragma unroll
    for (int j = 0; j < LPL; j++) { const int l = rank * LPL + j; if (l < len) nb.nw_off[l] = run; run += nn[j]; }
    if (rank == 31) {
      const uint32_t base = sh.ncells;
      uint32_t off = incl;
      if ((uint64_t)base + off > S.cap_cells) { a_or(&sh.status, GS_OVERFLOW); done = 0; off = 0; }
      nb.ch_base = base; nb.ch_len = (uint32_t)done;
      sh.ncells = base + off;
    }
  }
#else
  if (rank == 0) {   // place the levels: prefix sum, end of the border, chain runs
    const uint32_t base = sh.ncells;
    uint32_t off = 0;
    int done = len;
    const int bad = nb.nw_bad;
    for (int l = 0; l < len; l++) {
      if (l >= bad) { done = l; break; }
      const uint32_t nn = nb.nw_n[l];
      nb.nw_off[l] = off;
      off += nn;
      if (nn == 0) { done = l + 1; break; }
    }
    if ((uint64_t)base + off > S.cap_cells) { a_or(&sh.status, GS_OVERFLOW); done = 0; off = 0; }
    nb.ch_base = base; nb.ch_len = (uint32_t)done;
    sh.ncells = base + off;
  }
#endif
  G::sync();
  const int done = (int)ld_sh(&nb.ch_len);
  const uint32_t base = ld_sh(&nb.ch_base);
  for (int l = rank; l < done; l += gsz) {
    const uint32_t level = (uint32_t)(c + l);
    // everything the level needs from shared memory, read once into registers
    uint64_t st[NARROW_W], ps[NARROW_W];
    uint32_t av[NARROW_W], ev[NARROW_W], c0[NARROW_W];
    bool al[NARROW_W], pal[NARROW_W];
#pragma unroll
    for (int t = 0; t < NARROW_W; t++) {
      const bool in = t < w;
      const int fail = in ? nb.nw_fail[t] : 0, born = in ? nb.nw_born[t] : 0;
      al[t] = in && l < fail && l >= born;
      pal[t] = in && (l == 0 ? born == 0 : (l - 1 < fail && l - 1 >= born));
      st[t] = nb.ch_s[l][t];
      ps[t] = l ? nb.ch_s[l - 1][in && l == born && born > 0 ? nb.nw_par[t] : t] : nb.nw_x0[t];   // a new chain descends from its parent's state
      av[t] = nb.ch_a[l][t]; ev[t] = nb.ch_e[l][t]; c0[t] = nb.nw_cnt0[t];
    }
    const uint32_t off_l = nb.nw_off[l], off_p = l ? nb.nw_off[l - 1] : 0u, nn = nb.nw_n[l];
    // a chain's cell = the lowest alive chain with the same state; cells are laid out in that order
    uint32_t rep[NARROW_W], prep[NARROW_W];
#pragma unroll
    for (int t = 0; t < NARROW_W; t++) {
      rep[t] = (uint32_t)t; prep[t] = (uint32_t)t;
#pragma unroll
      for (int u = NARROW_W - 1; u >= 0; u--) {
        if (u >= t) continue;
        if (al[u] && st[u] == st[t]) rep[t] = (uint32_t)u;
        if (pal[u] && ps[u] == ps[t]) prep[t] = (uint32_t)u;
      }
    }
    uint32_t ord = 0;
#pragma unroll
    for (int t = 0; t < NARROW_W; t++) {
      if (!al[t] || rep[t] != (uint32_t)t) continue;
      const uint32_t idx = base + off_l + ord++;
      const uint64_t sx = st[t];
      const uint32_t a = av[t], en = ev[t], strand = (a >> 8) & 1;
      uint32_t cnt = 0, l0 = NONE32, l1 = NONE32, l2 = NONE32, l3 = NONE32;
      bool q3 = false;
#pragma unroll
      for (int v = 0; v < NARROW_W; v++) {
        if (!al[v]) continue;
        if (st[v] != sx) { if (ev[v] == en) q3 = true; continue; }   // both strands of one node on one level (Q3)
        cnt += c0[v];   // <= 4 * MAX_PATHS < 2^32
        // the cell of chain v on the level below, and its position among that level's distinct states
        uint32_t pord = 0;
#pragma unroll
        for (int u = 0; u < NARROW_W; u++) if ((uint32_t)u < prep[v] && pal[u] && prep[u] == (uint32_t)u) pord++;
        const uint32_t pidx = l ? base + off_p + pord : lo + prep[v];
        const uint32_t pb = (uint32_t)(ps[v] >> (2 * (k - 1))) & 3u;
        l0 = pb == 0 ? pidx : l0; l1 = pb == 1 ? pidx : l1; l2 = pb == 2 ? pidx : l2; l3 = pb == 3 ? pidx : l3;
      }
      S.ck[idx] = sx; S.cc[idx] = cnt; S.cn[idx] = en;
      *(uint4 *)(S.cp + 4 * (size_t)idx) = uint4{l0, l1, l2, l3};
      S.cflags[idx] = (a & 0xFF) << CF_OMASK_SHIFT;   // the scout emits the oriented out-mask
      if (q3) a_or(&sh.status, GS_ORDER_DEPENDENT);
      a_max(&S.ent[en].fstamp[strand], level + 1);
      S.ent[en].fcell[strand] = idx;
      if ((a & (1u << 17)) && (int)level >= sh.tlo) {
        for (int j = 0; j <= gd.rmf; j++)
          if ((S.rk[j] & FLANK_KMER) == sx) S.thit[(size_t)j * sh.trow + (level - (uint32_t)sh.tlo)] = idx + 1;
      }
    }
    S.lstart[level + 1] = base + off_l + nn;
  }
  G::sync();
  if (rank == 0) { sh.prof[2] += now_cycles() - tp2; sh.prof[5] += 1; sh.prof[6] += (unsigned long long)done; }
  return done;
}

// The node of a marked (Phase B) cell remembers the highest level it is marked on (bits 8.. of its flag word; the low
// bits are not read any more once the backward sweep runs).  A node marked on two levels makes the path subgraph a
// general graph (vertex identity is the node, Gap2Seq.cpp:1222-1297); the pack phase sees that as a marked cell whose
// level differs from its node's.  No value is returned, so the sweep never waits for the atomic.
// Returns whether the node carried a mark of another level already: the sweep runs top down, so such a mark comes
// from a higher level and the node is a vertex of two levels (general subgraph).  Callers collect the answers of
// several cells before looking at them, so that the atomics are in flight together.
G2S_D bool mark_node_level(const Slice &S, uint32_t en, int level) {
  const uint32_t old = a_max(&S.ent[en].flags, ((uint32_t)(level + 1) << 8) | 0xFFu);
  return (old >> 8) != 0 && (old >> 8) != (uint32_t)(level + 1);
}

G2S_D uint32_t popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__popc(x);
#else
  return (uint32_t)__builtin_popcount(x);
#endif
}

G2S_D uint32_t low_bit(uint32_t x) {   // index of the lowest set bit, x != 0
#ifdef __CUDA_ARCH__
  return (uint32_t)__ffs((int)x) - 1u;
#else
  return (uint32_t)__builtin_ctz(x);
#endif
}

// nibble i of M = set of cells of the level below that are predecessors of cell i; returns the union over the cells in v
G2S_D uint32_t flow_down(uint32_t M, uint32_t v) {
  const uint32_t s = (v | (v << 3) | (v << 6) | (v << 9)) & 0x1111u;   // bit i of v -> bit 4i
  const uint32_t sel = M & (s * 15u);
  const uint32_t r = sel | (sel >> 8);
  return (r | (r >> 4)) & 15u;
}

// Backward marking of up to CHAIN_LEN consecutive narrow levels d, d-1, ... (1..NARROW_W cells each, all above lmf,
// so none of their cells is a start cell).  One thread per level gathers the level's flags and predecessor links
// (all loads of a batch are in flight together), one thread then pushes the marks down the batch in shared memory
// (a level costs a few bit operations; a batch of pure chain levels only ORs), and the marks, pack indices and
// per-level counts are written back by one thread per level without waiting for any atomic.  Marks reach the level
// below the batch through the cell flags, exactly like on the level-by-level path.  Returns the number of levels
// done (0: level d is wide or empty).
template <class G>
G2S_D int narrow_backward(const Slice &S, GShared &sh, NarrowBuf &nb, int d, int lmf) {
  const int rank = G::rank(), gsz = G::size();
  const int maxlen = d - lmf < CHAIN_LEN ? d - lmf : CHAIN_LEN;   // levels d .. d-maxlen+1 are all > lmf
  if (rank == 0) { nb.bk_len = maxlen; nb.nw_stop = 0; }
  G::sync();
  // geometry: nw_off[t] = first cell of batch level t (= level d - t), nw_n[t] = its width
  for (int t = rank; t <= maxlen; t += gsz) {
    const uint32_t lo = ld_cg(&S.lstart[d - t]);
    nb.nw_off[t] = lo;
    if (t < maxlen) {
      const uint32_t n = ld_cg(&S.lstart[d + 1 - t]) - lo;
      nb.nw_n[t] = n;
      if (n == 0 || n > (uint32_t)NARROW_W) a_mini(&nb.bk_len, t);
    }
  }
  G::sync();
  const int len = ld_sh(&nb.bk_len);
  if (len == 0) return 0;
  // gather, one level per thread: own marks, predecessor sets relative to the level below, nodes
  for (int t = rank; t < len; t += gsz) {
    const uint32_t lo = nb.nw_off[t], n = nb.nw_n[t], below = nb.nw_off[t + 1];
    uint32_t f[NARROW_W], e[NARROW_W];
    uint4 p[NARROW_W];
#pragma unroll
    for (int i = 0; i < NARROW_W; i++) {
      f[i] = 0; e[i] = 0; p[i] = uint4{NONE32, NONE32, NONE32, NONE32};
      if ((uint32_t)i < n) {
        f[i] = ld_cg(&S.cflags[lo + i]);
        e[i] = ld_cg(&S.cn[lo + i]);
        p[i] = ld_cg((const uint4 *)(S.cp + 4 * (size_t)(lo + i)));
      }
    }
    uint32_t own = 0, M = 0, npw = 0;
#pragma unroll
    for (int i = 0; i < NARROW_W; i++) {
      if (f[i] & CF_B) own |= 1u << i;
      if (f[i] & CF_T) own |= 16u << i;
      const uint32_t pp[4] = {p[i].x, p[i].y, p[i].z, p[i].w};
      uint32_t pm = 0, np = 0;
#pragma unroll
      for (int z = 0; z < 4; z++)
        if (pp[z] != NONE32) { np++; const uint32_t rel = pp[z] - below; if (rel < (uint32_t)NARROW_W) pm |= 1u << rel; }
      M |= pm << (4 * i);
      npw |= np << (4 * i);
      nb.ch_e[t][i] = e[i];
      if (t == len - 1) { nb.bk_p[i][0] = pp[0]; nb.bk_p[i][1] = pp[1]; nb.bk_p[i][2] = pp[2]; nb.bk_p[i][3] = pp[3]; }
    }
    nb.ch_a[t][0] = own | (M << 8);
    nb.ch_a[t][1] = npw;
    if (M != 1u || n != 1u) a_or(&nb.nw_stop, 1u);   // not a pure chain level
  }
  G::sync();
#if defined(G2S_DEVICE_BUILD) && !defined(G2S_SERIAL_BACK)
  // Push the marks down the batch.  Level t maps the marks x that reach it to F_t(own_t | x) on the level below, F_t
  // = union of predecessor sets (flow_down).  Such maps compose - (A, c): x -> c | A x with a 4 x 4 boolean matrix A,
  // applied to the B and the T nibble alike - so the group's first warp scans them: every lane composes its
  // consecutive levels, one warp scan yields the marks entering each lane's levels, and the lanes finish their levels.
  if (rank < 32) {
    const unsigned FULL = 0xffffffffu;
    const int LPL = (CHAIN_LEN + 31) / 32;
    uint32_t own[LPL], Mx[LPL];
    uint32_t A = 0x8421u, cB = 0, cT = 0;   // identity, no marks
#pragma unroll
    for (int j = 0; j < LPL; j++) {
      const int t = rank * LPL + j;
      own[j] = 0; Mx[j] = 0x8421u;
      if (t < len) { const uint32_t w0 = nb.ch_a[t][0]; own[j] = w0 & 0xFFu; Mx[j] = w0 >> 8; }
      cB = flow_down(Mx[j], (own[j] & 15u) | cB);
      cT = flow_down(Mx[j], (own[j] >> 4) | cT);
      A = flow_down(Mx[j], A & 15u) | (flow_down(Mx[j], (A >> 4) & 15u) << 4) | (flow_down(Mx[j], (A >> 8) & 15u) << 8) |
          (flow_down(Mx[j], (A >> 12) & 15u) << 12);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {      // inclusive scan: (lanes before) then mine
      const uint32_t Ap = __shfl_up_sync(FULL, A, o), cBp = __shfl_up_sync(FULL, cB, o), cTp = __shfl_up_sync(FULL, cT, o);
      if (rank >= o) {
        cB |= flow_down(A, cBp);
        cT |= flow_down(A, cTp);
        A = flow_down(A, Ap & 15u) | (flow_down(A, (Ap >> 4) & 15u) << 4) | (flow_down(A, (Ap >> 8) & 15u) << 8) |
            (flow_down(A, (Ap >> 12) & 15u) << 12);
      }
    }
    uint32_t xB = __shfl_up_sync(FULL, cB, 1), xT = __shfl_up_sync(FULL, cT, 1);
    if (rank == 0) { xB = 0; xT = 0; }
    uint32_t mrk[LPL], cnt = 0;
#pragma unroll
    for (int j = 0; j < LPL; j++) {
      const uint32_t mB = (own[j] & 15u) | xB, mT = (own[j] >> 4) | xT;
      mrk[j] = rank * LPL + j < len ? mB | (mT << 4) : 0u;   // levels past the batch hold no cells
      cnt += popc32((mrk[j] | (mrk[j] >> 4)) & 15u);
      xB = flow_down(Mx[j], mB); xT = flow_down(Mx[j], mT);
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (rank >= o) incl += v; }
    uint32_t run = incl - cnt;
#pragma unroll
    for (int j = 0; j < LPL; j++) {
      const int t = rank * LPL + j;
      if (t < len) { nb.nw_run[t] = mrk[j]; nb.nw_poff[t] = run; }
      run += popc32((mrk[j] | (mrk[j] >> 4)) & 15u);
    }
    if (rank == 31) { nb.ch_base = sh.npack; sh.npack += incl; }
  }
#else
  if (rank == 0) {   // push the marks down the batch
    uint32_t x = 0, npk = 0;
    if (nb.nw_stop == 0) {
      for (int t = 0; t < len; t++) {
        x |= nb.ch_a[t][0] & 0x11u;
        nb.nw_run[t] = x;
        nb.nw_poff[t] = npk;
        npk += x ? 1u : 0u;
      }
    } else {
      for (int t = 0; t < len; t++) {
        const uint32_t w0 = nb.ch_a[t][0], M = w0 >> 8;
        const uint32_t m = (w0 & 0xFFu) | x;
        nb.nw_run[t] = m;
        nb.nw_poff[t] = npk;
        npk += popc32((m | (m >> 4)) & 15u);
        x = flow_down(M, m & 15u) | (flow_down(M, m >> 4) << 4);
      }
    }
    nb.ch_base = sh.npack;
    sh.npack += npk;
  }
#endif
  G::sync();
  const uint32_t base = ld_sh(&nb.ch_base);
  uint32_t nbt = 0, ned = 0;
  bool two_levels = false;
  for (int t = rank; t < len; t += gsz) {
    const uint32_t m = nb.nw_run[t], mB = m & 15u, mT = m >> 4, mk = mB | mT;
    if (!mk) continue;
    const int level = d - t;
    const uint32_t lo = nb.nw_off[t], npw = nb.ch_a[t][1];
    if (mB) { S.nB[level] = popc32(mB); nbt += popc32(mB); }
    if (mT) { S.nT[level] = popc32(mT); S.tcell[level] = lo + low_bit(mT); }
    uint32_t pi = base + nb.nw_poff[t];
#pragma unroll
    for (int i = 0; i < NARROW_W; i++) {
      if (!((mk >> i) & 1)) continue;
      const uint32_t idx = lo + (uint32_t)i;
      const uint32_t bits = (((mB >> i) & 1) ? (uint32_t)CF_B : 0u) | (((mT >> i) & 1) ? (uint32_t)CF_T : 0u);
      a_or(&S.cflags[idx], bits);
      S.packidx[idx] = pi; S.packlist[pi] = idx; S.packlevel[pi] = (uint32_t)level;
      pi++;
      if (bits & CF_B) { two_levels |= mark_node_level(S, nb.ch_e[t][i], level); ned += (npw >> (4 * i)) & 15u; }
      if (t == len - 1) {
#pragma unroll
        for (int z = 0; z < 4; z++) { const uint32_t p = nb.bk_p[i][z]; if (p != NONE32) a_or(&S.cflags[p], bits); }
      }
    }
  }
  if (two_levels) a_or(&sh.status, GS_GENERAL);
  if (nbt) a_add(&sh.nB, nbt);
  if (ned) a_add(&sh.n_edges, ned);
  G2S_STAT(emul_stats().back_chain_levels += (unsigned long long)len);
  G::sync();
  return len;
}

// ------------------------------------------------------------------------------------------------ one gap
// A gap is processed in phases - init, right pass, forward pass, backward marking + pack - written as functions of
// the executing thread group so that the kernels can map them differently: the wide tier and the host emulation
// run them one after the other on one group; the thin tier runs the right pass on a helper warp concurrently with
// the part of the forward pass that does not need the right set yet (levels below g/2 + e/2 + lmf).
struct GapCtx {
  int k, g, e, lmf, rmf, right_half, Dmax, prune_from, tlo;
  size_t trow;
  bool all_paths, skip_confident;
  const char *lstr, *rstr, *rsstr;
  Table T;   // the gap's graph
};

G2S_D Table gap_table(const Table &one, const Table *tables, const uint32_t *graph_of, uint32_t gap_id) {
  return tables ? tables[graph_of[gap_id]] : one;
}

G2S_D GapCtx make_ctx(const FillParams &P, const GapDev &gd) {
  GapCtx cx;
  cx.T = gap_table(P.table, P.tables, P.graph_of, gd.id);
  cx.k = P.table.k;
  cx.g = gd.gap_len; cx.e = gd.gap_err; cx.lmf = gd.lmf; cx.rmf = gd.rmf;
  cx.all_paths = (gd.flags & 2) != 0; cx.skip_confident = (gd.flags & 1) != 0;
  cx.right_half = cx.rmf + (cx.g + cx.e + 1) / 2;           // rmf + ceil((g+e)/2), Gap2Seq.cpp:862
  cx.Dmax = cx.g + cx.e + cx.lmf + cx.rmf;                   // right_half + left_half, Gap2Seq.cpp:863
  cx.prune_from = cx.g / 2 + cx.e / 2 + cx.lmf;              // two integer divisions, Gap2Seq.cpp:1050
  cx.lstr = P.flanks + gd.flank_off; cx.rstr = cx.lstr + cx.k + cx.lmf; cx.rsstr = P.flanks + gd.rstart_off;
  cx.tlo = cx.g + cx.lmf - cx.e > 0 ? cx.g + cx.lmf - cx.e : 0;
  cx.trow = (size_t)(cx.Dmax - cx.tlo + 1);
  return cx;
}

// the forward pass calls wait_right() exactly once, before the first level that consults the right set
struct NoWait {
  static constexpr bool kConcurrent = false;
  static G2S_D void wait_right() {}
};

// ---- init: clear the node table, thit and the per-level counters; parse the flanks; flag the right-flank targets
template <class G>
G2S_D void gap_init(const FillParams &P, const GapCtx &cx, const Slice &S, GShared &sh) {
  const Table &T = cx.T;
  const int k = cx.k, rank = G::rank(), gsz = G::size();
  uint4 *hz = (uint4 *)S.ent;
  const uint4 z0 = {0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u}, z1 = {0u, 0u, 0u, 0u};
  for (uint32_t i = rank; i < S.hcap; i += gsz) { hz[2 * i] = z0; hz[2 * i + 1] = z1; }
  const size_t nth = cx.trow * (size_t)(cx.rmf + 1);
  for (size_t i = rank; i < nth; i += gsz) S.thit[i] = 0;
  for (int i = rank; i <= cx.Dmax + 1; i += gsz) { S.nB[i] = 0; S.nT[i] = 0; S.lstart[i] = 0; }
  for (int d = rank; d <= cx.lmf; d += gsz) {
    uint64_t x = kmer_from_ascii(cx.lstr + d, k);
    uint32_t adj;
    Canon c = canonical(x, k);
    S.lk[d] = x | (table_find(T, c.key, &adj) >= 0 ? FLANK_PRESENT : 0);
  }
  for (int j = rank; j <= cx.rmf; j += gsz) {
    uint64_t x = kmer_from_ascii(cx.rstr + j, k);
    uint32_t adj;
    Canon c = canonical(x, k);
    S.rk[j] = x | (table_find(T, c.key, &adj) >= 0 ? FLANK_PRESENT : 0);
  }
  if (rank == 0) {
    sh.tlo = cx.tlo; sh.trow = (uint32_t)cx.trow;
    for (int q = 0; q < 8; q++) sh.prof[q] = 0;
    sh.table_lookups = (uint32_t)(cx.lmf + cx.rmf + 2); sh.node_ops = 0;
    sh.ncells = 0; sh.nnodes = 0; sh.status = 0;
    sh.nbuf[0].nnext = 0; sh.nbuf[1].nnext = 0;
    sh.npack = 0; sh.nB = 0; sh.n_edges = 0; sh.n_src = 0; sh.n_sink = 0;
    sh.max_src_level = -1; sh.min_sink_level = 0x7fffffff; sh.max_seed_level = -1;
    sh.cells_right = 0; sh.n_plen = 0; sh.plen[0] = sh.plen[1] = -1; sh.reached_fuz = 0; sh.count_fwd = 0;
    sh.found_level = 0x7fffffff; sh.count_all = 0; sh.pack_off = 0;
    sh.levels_r = 0; sh.levels_f = 0; sh.top_level = 0; sh.fwd_overflow = 0; sh.cyc_right = 0;
  }
  G::sync();
  // the right-flank targets: their node entries exist from the start, flagged TARGET, so that the forward pass
  // records a hit whenever it creates a cell of such a node (Gap2Seq.cpp:1119)
  for (int j = rank; j <= cx.rmf; j += gsz) {
    const uint64_t x = S.rk[j];
    if (x & FLANK_PRESENT) {
      const Canon c = canonical(x & FLANK_KMER, k);
      const uint32_t en = node_insert(S, sh, c.key);
      if (en != NONE32) a_or(&S.ent[en].flags, NF_TARGET);
    }
  }
  G::sync();
}

// ---- Phase R: level-synchronous backward walk from the staggered right starts (Gap2Seq.cpp:871-982).
// Only node membership in the right set is read later (Gap2Seq.cpp:1050): nodes get NF_IN_RS, cells are counted.
template <class G>
G2S_D void gap_right(const FillParams &P, const GapCtx &cx, const Slice &S, GShared &sh, NarrowBuf &nb) {
  const Table &T = cx.T;
  const int k = cx.k, rank = G::rank(), gsz = G::size(), rmf = cx.rmf, right_half = cx.right_half;
  const unsigned long long t0 = now_cycles();
  int levels_r = 0;
  uint32_t n = 0, my_lookups = 0;   // general-path successors: one table lookup + one node operation each
  int cur = 0;
  if (rank == 0) {
    uint64_t x = kmer_from_ascii(cx.rsstr + rmf, k);
    Canon c = canonical(x, k);
    uint32_t adj = 0;
    if (table_find(T, c.key, &adj) >= 0) {
      uint32_t en = node_insert(S, sh, c.key);
      if (en != NONE32) {
        a_max(&S.ent[en].rstamp, 2 + c.strand);
        a_or(&S.ent[en].flags, NF_IN_RS);
        S.rbk[0][0] = x; S.rbm[0][0] = (unsigned char)in_mask(adj, c.strand);
        nb.nnext = 1; sh.cells_right = 1;
      }
    }
  }
  G::sync();
  n = ld_sh(&nb.nnext);
  int prepared_for = -1;
  for (int c = 1; c <= right_half; c++) {
    if (n == 0 && c > rmf) break;
    levels_r = c;
    if (n >= 1 && n <= (uint32_t)NARROW_W && c > rmf) {   // narrow border, no start left to inject
      uint32_t n2 = n;
      const int adv = narrow_right<G>(T, S, sh, nb, cur, c, right_half, (int)n, &n2);
      if (group_overflow<G>(sh)) break;
      if (adv > 0) { n = n2; levels_r = c + adv - 1; c += adv - 1; continue; }
    }
    uint32_t *cnt = &nb.rnext[(uint32_t)c % 3u];
    if (prepared_for != c) {   // first general level, or a narrow batch came before: this level's counter is stale
      G::sync();
      if (rank == 0) *cnt = 0;
      G::sync();
    }
    if (rank == 0) nb.rnext[(uint32_t)(c + 1) % 3u] = 0;   // the next level's counter: it was last read two barriers ago
    prepared_for = c + 1;
    const uint64_t *fk = S.rbk[cur]; const unsigned char *fm = S.rbm[cur];
    uint64_t *nk = S.rbk[cur ^ 1]; unsigned char *nm = S.rbm[cur ^ 1];
#ifdef G2S_DEVICE_BUILD
    {   // pair step (see pair_scan)
      const uint32_t lane = (uint32_t)rank & 31u, warp = (uint32_t)rank >> 5, nwarps = (uint32_t)gsz >> 5;
      const uint32_t C = n <= 8u * nwarps ? 8u : n <= 16u * nwarps ? 16u : 32u;
      for (uint32_t base = 0; base < n; base += C * nwarps) {
        const uint32_t e0 = base + warp * C;
        uint64_t x = 0; uint32_t im = 0;
        if (lane < C && e0 + lane < n) { x = fk[e0 + lane]; im = fm[e0 + lane] & 15u; }
        const uint32_t nnodes_seen = ld_sh(&sh.nnodes);
        my_lookups += popc32(im);
        const PairSrc ps = pair_scan(popc32(im));
        for (uint32_t it = 0; it * 32u < ps.total; it++) {
          const uint32_t q = it * 32u + lane;
          const uint32_t src = pair_owner(ps, q, C);
          const uint64_t sx = shfl64(x, src);
          const uint32_t sim = __shfl_sync(0xffffffffu, im, src), sex = __shfl_sync(0xffffffffu, ps.excl, src);
          bool act = q < ps.total, win = false;
          uint64_t p = 0; uint32_t adj = 0, strand = 0;
          if (act) {
            p = pred_kmer(sx, nth_bit(sim, q - sex), T);
            const Canon cp = canonical(p, k);
            strand = cp.strand;
            uint32_t nf = 0;
            const uint32_t en = table_and_node<true>(T, S, sh, cp.key, nnodes_seen, &adj, &nf);
            if (en == NONE32) a_or(cnt, 0x80000000u);
            if (en != NONE32) {
              if (!(nf & NF_IN_RS)) a_or(&S.ent[en].flags, NF_IN_RS);
              const uint32_t stamp = (uint32_t)(c + 1) * 2 + strand;
              const uint32_t prev = a_max(&S.ent[en].rstamp, stamp);
              if ((prev >> 1) < (uint32_t)(c + 1)) win = true;
              else if ((prev & 1) != strand) a_or(&sh.status, GS_ORDER_DEPENDENT);   // both strands on one level
            }
          }
          const uint32_t idx = warp_reserve(cnt, win);
          if (win) {
            if (idx >= S.cap_cells) a_or(&sh.status, GS_OVERFLOW);
            else { nk[idx] = p; nm[idx] = (unsigned char)in_mask(adj, strand); }
          }
        }
      }
    }
#else
    for (uint32_t i = rank; i < n; i += gsz) {
      const uint64_t x = fk[i];
      uint32_t im = fm[i];
      my_lookups += popc32(im & 15u);
      for (uint32_t nt = 0; nt < 4; nt++) {
        if (!((im >> nt) & 1)) continue;
        uint64_t p = pred_kmer(x, nt, T);
        Canon cp = canonical(p, k);
        uint32_t adj = 0;
        if (table_find(T, cp.key, &adj) < 0) { a_or(&sh.status, GS_INTERNAL); continue; }
        uint32_t en = node_insert(S, sh, cp.key);
        if (en == NONE32) { a_or(cnt, 0x80000000u); continue; }
        a_or(&S.ent[en].flags, NF_IN_RS);
        uint32_t stamp = (uint32_t)(c + 1) * 2 + cp.strand;
        uint32_t prev = a_max(&S.ent[en].rstamp, stamp);
        if ((prev >> 1) < (uint32_t)(c + 1)) {
          uint32_t idx = a_inc(cnt);
          if (idx >= S.cap_cells) { a_or(&sh.status, GS_OVERFLOW); continue; }
          nk[idx] = p; nm[idx] = (unsigned char)in_mask(adj, cp.strand);
        } else if ((prev & 1) != cp.strand) {
          a_or(&sh.status, GS_ORDER_DEPENDENT);   // both strands on one level
        }
      }
    }
#endif
    G::sync();
    if (c <= rmf) {
      if (rank == 0) {   // next staggered right start (Gap2Seq.cpp:953-968)
        uint64_t x = kmer_from_ascii(cx.rsstr + (rmf - c), k);
        Canon cs = canonical(x, k);
        uint32_t adj = 0;
        if (table_find(T, cs.key, &adj) >= 0) {
          uint32_t en = node_insert(S, sh, cs.key);
          if (en == NONE32) a_or(cnt, 0x80000000u);
          if (en != NONE32) {
            a_or(&S.ent[en].flags, NF_IN_RS);
            uint32_t stamp = (uint32_t)(c + 1) * 2 + cs.strand;
            uint32_t prev = a_max(&S.ent[en].rstamp, stamp);
            if ((prev >> 1) < (uint32_t)(c + 1)) {
              uint32_t idx = a_add(cnt, 1);
              if (idx >= S.cap_cells) a_or(&sh.status, GS_OVERFLOW);
              else { nk[idx] = x; nm[idx] = (unsigned char)in_mask(adj, cs.strand); }
            } else if ((prev & 1) != cs.strand) {
              a_or(&sh.status, GS_ORDER_DEPENDENT);
            }
          }
        }
      }
      G::sync();
    }
    const uint32_t raw = ld_sh(cnt);   // stable: the next append to this counter is three levels away
    n = raw & 0x7FFFFFFFu;
    if ((raw >> 31) != 0 || n > S.cap_cells) {   // this pass outgrew its slice (the bit is in the status word already)
      if (rank == 0) a_or(&sh.status, GS_OVERFLOW);
      break;
    }
    if (rank == 0) sh.cells_right += n;
    cur ^= 1;
  }
  if (my_lookups) { a_add(&sh.table_lookups, my_lookups); a_add(&sh.node_ops, my_lookups); }
  G::sync();
  if (rank == 0) { sh.levels_r = levels_r; sh.cyc_right = now_cycles() - t0; }
  G::sync();
}

// ---- Phase F: forward walk counts by exact length (Gap2Seq.cpp:984-1167)
template <class G, class W>
G2S_D void gap_forward(const FillParams &P, const GapDev &gd, const GapCtx &cx, const Slice &S, GShared &sh, NarrowBuf &nb) {
  const Table &T = cx.T;
  const int k = cx.k, rank = G::rank(), gsz = G::size();
  const int g = cx.g, lmf = cx.lmf, rmf = cx.rmf, Dmax = cx.Dmax, prune_from = cx.prune_from;
  const bool all_paths = cx.all_paths;
  int top_level = 0;       // last level that was expanded into
  int levels_f = 0;
  uint32_t my_lookups = 0;
  bool waited = !W::kConcurrent;
  bool overflow = group_overflow<G>(sh);
  if (!overflow) {
    if (rank == 0) {
      uint64_t x = S.lk[0];
      if (x & FLANK_PRESENT) {
        x &= FLANK_KMER;
        Canon c = canonical(x, k);
        uint32_t adj = 0;
        table_find(T, c.key, &adj);
        uint32_t nf = 0;
        uint32_t en = node_insert(S, sh, c.key, &nf);
        if (en != NONE32) {
          S.ent[en].fstamp[c.strand] = 1;
          new_cell<G>(S, sh, gd, x, en, c.strand, adj, 0, 1, nf);
        }
      }
      S.lstart[0] = 0;
    }
    G::sync();
    if (rank == 0) S.lstart[1] = sh.ncells;
    G::sync();

    bool found = false;
    int known_for = -1; uint32_t known_lo = 0, known_hi = 0;
    for (int c = 1; c <= Dmax; c++) {
      if (!waited && c >= prune_from) { W::wait_right(); waited = true; }
      // until the right set is complete a look-ahead batch must end below the first pruned level
      const int dlim = waited ? Dmax : (prune_from - 1 < Dmax ? prune_from - 1 : Dmax);
      // the bounds of the border: handed over in registers by the general level before, else read back
      const uint32_t lo = known_for == c ? known_lo : ld_cg(&S.lstart[c - 1]), hi = known_for == c ? known_hi : ld_cg(&S.lstart[c]);
      const bool dead = (lo == hi) && c > lmf;   // empty border and no start left to inject
      if (hi > lo && hi - lo <= (uint32_t)NARROW_W) {   // narrow border (start levels included, see narrow_forward)
        const int adv = narrow_forward<G>(T, S, sh, nb, gd, lo, (int)(hi - lo), c, dlim, prune_from);
        if (group_overflow<G>(sh)) { overflow = true; break; }
        if (adv > 0) {
          const int last = c + adv - 1;
          top_level = levels_f = last;
          if (!found && last >= g + lmf + rmf) {
            const unsigned long long tf0 = now_cycles();
            found = found_scan<G>(S, gd, sh, c > g + lmf + rmf ? c : g + lmf + rmf, last);
            if (rank == 0) sh.prof[3] += now_cycles() - tf0;
            if (found && !all_paths) {   // -best-only: the table ends at the level of the first hit
              const int fl = ld_sh(&sh.found_level);
              G::sync();
              if (rank == 0) sh.ncells = S.lstart[fl + 1];
              top_level = levels_f = fl;
              G::sync();
              break;
            }
          }
          c = last;
          continue;
        }
      }
      const unsigned long long tg0 = now_cycles();
      if (rank == 0) sh.prof[7] += 1;
      if (!dead) {
        top_level = c;
        levels_f = c;
        // the cell counter and the status word as of the level's last reservation: read between the last round's two
        // barriers, when nobody writes them (the next level's first phase is a barrier away), so that the level needs no
        // third barrier for it
        uint32_t nc_last = 0, st_last = 0;
#ifdef G2S_DEVICE_BUILD
        {   // pair step (see pair_scan): C elements per warp, the smallest C that covers the border in one round
          const uint32_t lane = (uint32_t)rank & 31u, warp = (uint32_t)rank >> 5, nwarps = (uint32_t)gsz >> 5;
          const uint32_t wdt = hi - lo;
          const uint32_t C = wdt <= 8u * nwarps ? 8u : wdt <= 16u * nwarps ? 16u : 32u;
          const uint32_t lvl = (uint32_t)c + 1;
          for (uint32_t base = lo; base < hi; base += C * nwarps) {
            const uint32_t cell0 = base + warp * C;
            uint64_t x = 0; uint32_t om = 0, np = 0;
            if (lane < C && cell0 + lane < hi) {
              x = ld_cg(&S.ck[cell0 + lane]);
              om = (ld_cg(&S.cflags[cell0 + lane]) >> CF_OMASK_SHIFT) & 15u;
              np = clamp_paths(ld_cg(&S.cc[cell0 + lane]));
            }
            const uint32_t nnodes_seen = c < prune_from ? ld_sh(&sh.nnodes) : 0u;
            my_lookups += popc32(om);
            const PairSrc ps = pair_scan(popc32(om));
            // The pair that stamps a (node, strand) first creates the cell and writes it COMPLETE - its own walk count and
            // predecessor link included; only the pairs that find the stamp in place (two states converging on one
            // successor: rare) add theirs after the barrier.
            uint32_t rec_e[4], rec_i[4];
#pragma unroll
            for (uint32_t it = 0; it < 4; it++) {
              rec_e[it] = NONE32; rec_i[it] = 0;
              if (it * 32u < ps.total) {   // the same for the whole warp
                const uint32_t p = it * 32u + lane;
                const uint32_t src = pair_owner(ps, p, C);
                const uint64_t sx = shfl64(x, src);
                const uint32_t som = __shfl_sync(0xffffffffu, om, src), sex = __shfl_sync(0xffffffffu, ps.excl, src);
                const uint32_t snp = __shfl_sync(0xffffffffu, np, src);
                bool act = p < ps.total;
                uint32_t en = NONE32, adj = 0, nf = 0, strand = 0;
                uint64_t sk = 0;
                if (act) {
                  const uint32_t nt = nth_bit(som, p - sex);
                  sk = succ_kmer(sx, nt, T);
                  const Canon cs = canonical(sk, k);
                  strand = cs.strand;
                  if (c < prune_from) {
                    en = table_and_node<true>(T, S, sh, cs.key, nnodes_seen, &adj, &nf);
                  } else {   // meet in the middle: only nodes of the right set survive (Gap2Seq.cpp:1050)
                    en = table_and_node<false>(T, S, sh, cs.key, 0u, &adj, &nf);
                    if (en != NONE32 && !(nf & NF_IN_RS)) en = NONE32;
                  }
                  act = en != NONE32;
                }
                bool win = false;
                if (act) {
                  const uint32_t prev = a_max(&S.ent[en].fstamp[strand], lvl);
                  win = prev < lvl;
                  // Both strands of the node on this level?  The other strand's stamp is read AFTER this
                  // one is in place - the load's address is made to depend on the atomic's result - so of two pairs
                  // that stamp the two strands at least the later one sees the other.
                  uint32_t zero;
                  asm volatile("and.b32 %0, %1, 0;" : "=r"(zero) : "r"(prev));
                  if (ld_cg(&S.ent[en].fstamp[strand ^ 1] + zero) == lvl) a_or(&sh.status, GS_ORDER_DEPENDENT);
                }
                const uint32_t idx = warp_reserve(&sh.ncells, win);
                if (win) {
                  if (idx >= S.cap_cells) a_or(&sh.status, GS_OVERFLOW);
                  else new_cell_at(S, sh, gd, idx, sk, en, strand, adj, (uint32_t)c, snp, nf, (uint32_t)(sx >> (2 * (k - 1))) & 3u, cell0 + src);
                } else if (act) { rec_e[it] = en | (strand << 31); rec_i[it] = cell0 + src; }
              }
            }
            G::sync();
            if (base + C * nwarps >= hi) { nc_last = ld_sh(&sh.ncells); st_last = ld_sh(&sh.status); }
            if (!(ld_sh(&sh.status) & GS_OVERFLOW)) {
#pragma unroll
              for (uint32_t it = 0; it < 4; it++) {
                if (rec_e[it] == NONE32) continue;
                const uint32_t st = rec_e[it] >> 31, i = rec_i[it];
                const uint32_t idx = ld_cg(&S.ent[rec_e[it] & 0x7FFFFFFFu].fcell[st]);
                const uint32_t np2 = clamp_paths(ld_cg(&S.cc[i]));
                const uint32_t u0 = (uint32_t)(ld_cg(&S.ck[i]) >> (2 * (k - 1))) & 3u;
                a_add(&S.cc[idx], np2);
                S.cp[4 * idx + u0] = i;
              }
            }
          }
        }
#else
        for (uint32_t base = lo; base < hi; base += gsz) {
          const uint32_t i = base + rank;
          uint32_t cand_e[4];
          uint32_t cand_valid = 0, cand_strand = 0, np = 0, u0 = 0;
          if (i < hi) {
            const uint64_t x = ld_cg(&S.ck[i]);
            const uint32_t om = (ld_cg(&S.cflags[i]) >> CF_OMASK_SHIFT) & 15u;
            np = clamp_paths(ld_cg(&S.cc[i]));
            u0 = (uint32_t)(x >> (2 * (k - 1))) & 3u;
            my_lookups += popc32(om);
#pragma unroll
            for (uint32_t nt = 0; nt < 4; nt++) {
              cand_e[nt] = NONE32;
              if (!((om >> nt) & 1)) continue;
              const uint64_t s = succ_kmer(x, nt, T);
              const Canon cs = canonical(s, k);
              uint32_t adj = 0;
              if (table_find(T, cs.key, &adj) < 0) { a_or(&sh.status, GS_INTERNAL); continue; }
              uint32_t en, nf = 0;
              if (c < prune_from) {
                en = node_insert(S, sh, cs.key, &nf);
                if (en == NONE32) continue;
              } else {   // meet in the middle: only nodes of the right set survive (Gap2Seq.cpp:1050)
                en = node_find(S, cs.key, &nf);
                if (en == NONE32 || !(nf & NF_IN_RS)) continue;
              }
              const uint32_t prev = a_max(&S.ent[en].fstamp[cs.strand], (uint32_t)c + 1);
              if (prev < (uint32_t)c + 1) {
                if (new_cell<G>(S, sh, gd, s, en, cs.strand, adj, (uint32_t)c, 0, nf) == NONE32) continue;
              }
              cand_e[nt] = en;
              cand_valid |= 1u << nt;
              cand_strand |= cs.strand << nt;
            }
          }
          G::sync();
          if (cand_valid && !(ld_sh(&sh.status) & GS_OVERFLOW)) {
#pragma unroll
            for (uint32_t nt = 0; nt < 4; nt++) {
              if (!((cand_valid >> nt) & 1)) continue;
              const uint32_t st = (cand_strand >> nt) & 1;
              const NodeEnt *en = &S.ent[cand_e[nt]];
              const uint32_t idx = ld_cg(&en->fcell[st]);
              a_add(&S.cc[idx], np);
              S.cp[4 * idx + u0] = i;
              if (ld_cg(&en->fstamp[st ^ 1]) == (uint32_t)c + 1) a_or(&sh.status, GS_ORDER_DEPENDENT);
            }
          }
        }
        nc_last = sh.ncells; st_last = sh.status;
#endif
        G::sync();
        bool ovf;
        if (c <= lmf) {
          if (rank == 0 && c <= lmf) {   // staggered left start: count OVERWRITTEN with 1 (Gap2Seq.cpp:1082-1105)
            uint64_t x = S.lk[c];
            if (x & FLANK_PRESENT) {
              x &= FLANK_KMER;
              Canon cs = canonical(x, k);
              uint32_t adj = 0;
              table_find(T, cs.key, &adj);
              uint32_t nf = 0;
              uint32_t en = node_insert(S, sh, cs.key, &nf);
              if (en != NONE32) {
                if (S.ent[en].fstamp[cs.strand] == (uint32_t)c + 1) {
                  S.cc[S.ent[en].fcell[cs.strand]] = 1;
                } else {
                  S.ent[en].fstamp[cs.strand] = (uint32_t)c + 1;
                  new_cell<G>(S, sh, gd, x, en, cs.strand, adj, (uint32_t)c, 1, nf);
                }
                if (S.ent[en].fstamp[cs.strand ^ 1] == (uint32_t)c + 1) a_or(&sh.status, GS_ORDER_DEPENDENT);
              }
            }
          }
          G::sync();
          if (rank == 0) S.lstart[c + 1] = sh.ncells < S.cap_cells ? sh.ncells : S.cap_cells;
          ovf = group_overflow<G>(sh);
        } else {
          known_for = c + 1; known_lo = hi; known_hi = nc_last < S.cap_cells ? nc_last : S.cap_cells;
          if (rank == 0) S.lstart[c + 1] = known_hi;
          // (the thin tier's right pass runs on other warps at the same time and may set a bit at any moment: a vote there)
          ovf = W::kConcurrent ? group_overflow<G>(sh) : (st_last & GS_OVERFLOW) != 0;
        }
        if (ovf) { overflow = true; break; }
      } else {
        // the border died: no cell can appear any more; finish the found check for the remaining levels in
        // parallel (one level per thread) and leave the level loop
        if (!found && Dmax >= g + lmf + rmf) found = found_scan<G>(S, gd, sh, c > g + lmf + rmf ? c : g + lmf + rmf, Dmax);
        break;
      }
      // found check (Gap2Seq.cpp:1108-1154): one target per thread, then thread 0 takes the first in j order
      if (!found && c >= g + lmf + rmf) {
        G::sync();
        const int err = c - g - lmf - rmf;
        for (int j = rank; j <= rmf; j += gsz) {
          int len1 = g + lmf + j + err, len2 = g + lmf + j - err;
          sh.fc1[j] = thit_count(S, sh, j, len1);
          sh.fc2[j] = (len2 != len1 && len2 >= 0) ? thit_count(S, sh, j, len2) : 0;
        }
        G::sync();
        if (rank == 0) {
          for (int j = 0; j <= rmf; j++) {
            if (sh.fc1[j] | sh.fc2[j]) {
              uint64_t tot = 0; uint32_t n = 0;
              if (sh.fc1[j]) { tot += sh.fc1[j]; sh.plen[n++] = g + lmf + j + err; }
              if (sh.fc2[j]) { tot += sh.fc2[j]; sh.plen[n++] = g + lmf + j - err; }
              sh.n_plen = n; sh.reached_fuz = j;
              sh.count_fwd = tot > MAX_PATHS ? MAX_PATHS : (uint32_t)tot;
              sh.found_level = c;
              break;
            }
          }
        }
        G::sync();
        found = ld_sh(&sh.n_plen) != 0;
        if (found && !all_paths) break;   // -best-only stops at the first hit (Gap2Seq.cpp:1156-1158)
      } else if (c <= lmf) {
        G::sync();
      }
      if (rank == 0) sh.prof[4] += now_cycles() - tg0;
    }
  }
  if (!waited) W::wait_right();
  if (my_lookups) { a_add(&sh.table_lookups, my_lookups); a_add(&sh.node_ops, my_lookups); }
  G::sync();
  if (rank == 0) { sh.top_level = top_level; sh.levels_f = levels_f; sh.fwd_overflow = overflow ? 1 : 0; }
  G::sync();
}

// ---- Phase B + traceback trellis: backward marking through the predecessor links (Gap2Seq.cpp:1173-1312), then
// the marked cells are appended to the output arena.  Returns the cycle stamp taken between the two.
template <class G>
G2S_D unsigned long long gap_backward_pack(const FillParams &P, const GapCtx &cx, const Slice &S, GShared &sh, NarrowBuf &nb) {
  const int k = cx.k, rank = G::rank(), gsz = G::size();
  const int g = cx.g, e = cx.e, lmf = cx.lmf, rmf = cx.rmf;
  const bool all_paths = cx.all_paths, skip_confident = cx.skip_confident;
  const int top_level = ld_sh(&sh.top_level);
  const bool overflow = group_overflow<G>(sh) || ld_sh(&sh.fwd_overflow) != 0;
  unsigned long long t_back = now_cycles();
  const uint32_t n_plen = ld_sh(&sh.n_plen);
  const uint32_t count_fwd = ld_sh(&sh.count_fwd);
  if (!overflow && count_fwd > 0 && n_plen > 0) {
    const int J = ld_sh(&sh.reached_fuz);
    const int p0 = ld_sh(&sh.plen[0]), p1 = n_plen > 1 ? ld_sh(&sh.plen[1]) : -1;
    // seeds
    if (!skip_confident) {
      if (all_paths) {
        if (rmf >= 1) {
          const int jt = rmf - 1;   // Gap2Seq.cpp:1197-1206: effectively only target rmf-1
          int dlo = lmf + g - e; if (dlo < 0) dlo = 0;
          for (int d = dlo + rank; d <= top_level; d += gsz) {
            uint32_t t = thit_get(S, sh, jt, d);
            if (!t) continue;
            a_or(&S.cflags[t - 1], CF_B | CF_SEED);
            a_add64(&sh.count_all, clamp_paths(ld_cg(&S.cc[t - 1])));
            a_add(&sh.n_sink, 1);
            a_mini(&sh.min_sink_level, d);
            a_maxi(&sh.max_seed_level, d);
          }
        }
      } else if (rank == 0) {
        for (uint32_t q = 0; q < n_plen; q++) {
          int d = q ? p1 : p0;
          uint32_t t = thit_get(S, sh, J, d);
          if (t) { a_or(&S.cflags[t - 1], CF_B | CF_SEED); sh.n_sink++; if (d < sh.min_sink_level) sh.min_sink_level = d; a_maxi(&sh.max_seed_level, d); }
        }
      }
    }
    if (rank == 0) {
      for (uint32_t q = 0; q < n_plen; q++) {
        uint32_t t = thit_get(S, sh, J, q ? p1 : p0);
        if (t) { a_or(&S.cflags[t - 1], CF_T); a_maxi(&sh.max_seed_level, q ? p1 : p0); }
      }
    }
    G::sync();

    // nothing above the highest seed can be marked
    const int sweep_top = ld_sh(&sh.max_seed_level) < top_level ? ld_sh(&sh.max_seed_level) : top_level;
    // the bounds of a level are read one level ahead, so that the sweep's chain of dependent reads starts at the flags;
    // a level wider than the narrow mode's (or empty) does not even ask for a batch (narrow_backward would decline it
    // after two barriers)
    uint32_t b_hi = sweep_top >= 0 ? ld_cg(&S.lstart[sweep_top + 1]) : 0u, b_lo = sweep_top >= 0 ? ld_cg(&S.lstart[sweep_top]) : 0u;
    for (int d = sweep_top; d >= 0; d--) {
      const uint32_t lo = b_lo, hi = b_hi;
      const uint32_t lo_below = d > 0 ? ld_cg(&S.lstart[d - 1]) : 0u;
      if (d > lmf && hi - lo >= 1u && hi - lo <= (uint32_t)NARROW_W) {   // batches of narrow levels (no start cells up there)
        const int adv = narrow_backward<G>(S, sh, nb, d, lmf);
        if (adv > 0) {
          d -= adv - 1;
          b_hi = ld_cg(&S.lstart[d]); b_lo = d > 0 ? ld_cg(&S.lstart[d - 1]) : 0u;   // bounds of level d - 1, the next one
          continue;
        }
      }
      b_hi = lo; b_lo = lo_below;
      G2S_STAT(emul_stats().back_general_levels++; emul_stats().width_hist[hi - lo > 7 ? 7 : hi - lo]++);
      const uint64_t lcanon = (d <= lmf) ? canonical(S.lk[d] & FLANK_KMER, k).key : KEY_EMPTY;
      uint32_t my_b = 0, my_t = 0, my_e = 0;   // this thread's share of the level's totals, added once per warp below
      bool two_levels = false;
      for (uint32_t i = lo + rank; i < hi; i += gsz) {
        const uint32_t f = ld_cg(&S.cflags[i]);
        if (!(f & (CF_B | CF_T))) continue;
        const uint32_t pi = a_inc(&sh.npack);
        S.packidx[i] = pi; S.packlist[pi] = i; S.packlevel[pi] = (uint32_t)d;
        const uint32_t en = S.cn[i];
        const bool start = d <= lmf && ld_cg(&S.ent[en].key) == lcanon;
        uint32_t p[4];
        p[0] = ld_cg(&S.cp[4 * i]); p[1] = ld_cg(&S.cp[4 * i + 1]); p[2] = ld_cg(&S.cp[4 * i + 2]); p[3] = ld_cg(&S.cp[4 * i + 3]);
        if (f & CF_B) {
          my_b++;
          two_levels |= mark_node_level(S, en, d);
          if (start) { a_add(&sh.n_src, 1); a_maxi(&sh.max_src_level, d); }
          else {
#pragma unroll
            for (int q = 0; q < 4; q++) if (p[q] != NONE32) { a_or(&S.cflags[p[q]], CF_B); my_e++; }
          }
        }
        if (f & CF_T) { my_t++; S.tcell[d] = i; }
        if ((f & CF_T) && !start) {
#pragma unroll
          for (int q = 0; q < 4; q++) if (p[q] != NONE32) a_or(&S.cflags[p[q]], CF_T);
        }
      }
      if (two_levels) a_or(&sh.status, GS_GENERAL);
      a_add_warp(&S.nB[d], my_b);
      a_add_warp(&sh.nB, my_b);
      a_add_warp(&sh.n_edges, my_e);
      a_add_warp(&S.nT[d], my_t);
      G::sync();
    }
    t_back = now_cycles();

    // ---- traceback on the device.  The reference's random traceback (Gap2Seq.cpp:1437-1517) walks the trellis of
    // CF_T cells from the accepted target down to a start cell and draws rand() at every step; when every level of
    // that trellis holds exactly one cell (and there is one path length) no draw can change anything: the fill is
    // that chain's last bases, upper case within k levels below a safe vertex (or the path's end), and the number of
    // draws is the chain's length.  The group writes the fill string itself, one level per thread, and the gap
    // needs neither pack records nor a host traceback - the host only advances rand() by the reported count.
    const int max_src_t = ld_sh(&sh.max_src_level), min_sink_t = ld_sh(&sh.min_sink_level);
    const bool sink_safe_t = ld_sh(&sh.n_sink) > 0;
    const bool general_now = group_status_any<G>(sh, GS_GENERAL);
    if (P.fills != nullptr && n_plen == 1 && (skip_confident || !general_now)) {
      const int d0 = p0;
      if (rank == 0) { sh.t_lo = 0x7fffffff; sh.t_cnt = 0; sh.t_bad = 0; sh.t_ok = 0; }
      G::sync();
      int lo_l = 0x7fffffff; uint32_t c1 = 0, bad = 0;
      for (int d = rank; d <= d0; d += gsz) {
        const uint32_t n = ld_cg(&S.nT[d]);
        if (n) lo_l = d < lo_l ? d : lo_l;
        c1 += n == 1 ? 1u : 0u;
        bad |= n > 1 ? 1u : 0u;
      }
      if (lo_l != 0x7fffffff) a_mini(&sh.t_lo, lo_l);
      if (c1) a_add(&sh.t_cnt, c1);
      if (bad) a_or(&sh.t_bad, 1u);
      G::sync();
      if (rank == 0) {
        const int d_end = sh.t_lo;
        bool ok = !sh.t_bad && d_end <= lmf && d_end <= d0 && sh.t_cnt == (uint32_t)(d0 - d_end + 1);
        if (ok) {   // the chain must end on a start cell (otherwise: "Unable to backtrace", left to the host)
          const uint32_t idx = ld_cg(&S.tcell[d_end]);
          ok = ld_cg(&S.ent[ld_cg(&S.cn[idx])].key) == canonical(S.lk[d_end] & FLANK_KMER, k).key;
        }
        if (ok) {
          const unsigned long long need = (unsigned long long)(d0 - d_end) + 1;
          const unsigned long long off = a_add64(P.fill_cursor, need);
          if (off + need > P.fill_cap) ok = false;
          else { sh.fill_off = off; sh.t_left_fuz = lmf - d_end; }
        }
        sh.t_ok = ok ? 1u : 0u;
      }
      G::sync();
      if (ld_sh(&sh.t_ok)) {
        const int d_end = ld_sh(&sh.t_lo);
        // four levels per thread and step: the dependent loads (cell -> node -> flags -> nB) of the four are in flight together
        const int TU = 4;
        for (int dbase = d_end + 1 + rank * TU; dbase <= d0; dbase += gsz * TU) {   // safe vertex? (the rule of the pack records below)
          uint32_t ci[TU], en[TU], lv[TU], nb1[TU];
          uint64_t kx[TU];
#pragma unroll
          for (int u = 0; u < TU; u++) ci[u] = dbase + u <= d0 ? ld_cg(&S.tcell[dbase + u]) : 0u;
#pragma unroll
          for (int u = 0; u < TU; u++) { en[u] = dbase + u <= d0 ? ld_cg(&S.cn[ci[u]]) : 0u; kx[u] = dbase + u <= d0 ? ld_cg(&S.ck[ci[u]]) : 0ull; }
#pragma unroll
          for (int u = 0; u < TU; u++) lv[u] = (dbase + u <= d0 && !skip_confident) ? ld_cg(&S.ent[en[u]].flags) >> 8 : 0u;
#pragma unroll
          for (int u = 0; u < TU; u++) nb1[u] = (dbase + u <= d0 && lv[u] != 0) ? ld_cg(&S.nB[lv[u] - 1]) : 0u;
#pragma unroll
          for (int u = 0; u < TU; u++) {
            if (dbase + u > d0) continue;
            bool safe = skip_confident;
            if (!safe) {
              if (lv[u] == 0) safe = sink_safe_t;
              else { const int bl = (int)lv[u] - 1; safe = nb1[u] == 1 && bl >= max_src_t && bl <= min_sink_t; }
            }
            // nT now carries the safe bit and the level's base: (safe ? 2 : 1) | base << 2
            S.nT[dbase + u] = (safe ? 2u : 1u) | ((uint32_t)(kx[u] & 3u) << 2);
          }
        }
        G::sync();
        char *outc = P.fills + ld_sh(&sh.fill_off);
        for (int d = d_end + 1 + rank; d <= d0; d += gsz) {
          const uint32_t me = ld_cg(&S.nT[d]);
          bool up = d0 - d < k || (me & 3u) == 2u;
          for (int z = d + 1; !up && z <= d0 && z < d + k; z++) up = (ld_cg(&S.nT[z]) & 3u) == 2u;
          const uint32_t b = me >> 2;
          const char ch = b == 0 ? 'A' : b == 1 ? 'C' : b == 2 ? 'T' : 'G';
          outc[d - 1 - d_end] = up ? ch : (char)(ch | 0x20);
        }
        if (rank == 0) { outc[d0 - d_end] = 0; sh.npack = 0; a_or(&sh.status, GS_TRACED); }
        G::sync();
      }
    }

    // ---- pack: append the marked cells to the output arena
    const uint32_t npack = ld_sh(&sh.npack);
    if (rank == 0) {
      unsigned long long off = a_add64(P.pack_cursor, (unsigned long long)npack);
      sh.pack_off = off;
      if (off + npack > P.pack_cap) a_or(&sh.status, GS_OUT_FULL);
    }
    G::sync();
    if (!(ld_sh(&sh.status) & GS_OUT_FULL)) {
      PackRec *outp = P.pack + ld_sh(&sh.pack_off);
      const int max_src = ld_sh(&sh.max_src_level), min_sink = ld_sh(&sh.min_sink_level);
      const bool sink_safe = ld_sh(&sh.n_sink) > 0;   // the sink lies on every source->sink path unless isolated (Q6)
      for (uint32_t pi = rank; pi < npack; pi += gsz) {
        const uint32_t i = S.packlist[pi];
        const uint32_t d = S.packlevel[pi];
        const uint32_t f = ld_cg(&S.cflags[i]);
        const uint32_t en = S.cn[i];
        const uint64_t x = S.ck[i];
        const bool start = (int)d <= lmf && S.ent[en].key == canonical(S.lk[d] & FLANK_KMER, k).key;
        PackRec r;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          uint32_t p = S.cp[4 * i + q];
          r.pred[q] = (p == NONE32 || start) ? NONE32 : S.packidx[p];
        }
        r.node = en;
        r.level = d;
        uint32_t meta = (f & 7u) | (start ? PM_START : 0) | ((uint32_t)(x & 3) << PM_BASE_SHIFT);
        if (!skip_confident) {
          // level-DAG case: a vertex is on every source->sink path iff it is alone on its level and no source
          // or sink edge jumps over that level.  Vertex identity is the node: a node marked on two
          // levels collapses them (general subgraph -> the host runs the SCC analysis and ignores these bits).
          const uint32_t lv = ld_cg(&S.ent[en].flags) >> 8;   // highest marked level of the node + 1, 0 = none
          if ((f & CF_B) && lv != d + 1) a_or(&sh.status, GS_GENERAL);
          bool safe;
          if (lv == 0) safe = sink_safe;   // node2boost default-inserts the sink
          else { int bl = (int)lv - 1; safe = S.nB[bl] == 1 && bl >= max_src && bl <= min_sink; }
          if (safe) meta |= PM_SAFE;
        }
        r.meta = meta;
        r.count = clamp_paths(S.cc[i]);
        outp[pi] = r;
      }
    }
  }
  G::sync();
  return t_back;
}

template <class G>
G2S_D void gap_finish(const Slice &S, GShared &sh, GapOut *out, const unsigned long long *t) {
  if (G::rank() == 0) {
    const bool overflow = sh.fwd_overflow != 0 || (sh.status & GS_OVERFLOW) != 0;
    const uint32_t ncells = sh.ncells < S.cap_cells ? sh.ncells : S.cap_cells;
    GapOut o;
    o.status = sh.status | GS_DONE;
    o.count_fwd = sh.count_fwd;
    o.count_all = sh.count_all > MAX_PATHS ? MAX_PATHS : (uint32_t)sh.count_all;
    o.n_plen = sh.n_plen; o.plen[0] = sh.plen[0]; o.plen[1] = sh.plen[1];
    o.reached_fuz = sh.reached_fuz;
    o.cells_right = sh.cells_right; o.cells_left = ncells; o.nodes = sh.nnodes;
    o.npack = sh.npack;
    o.start_pack[0] = o.start_pack[1] = NONE32;
    if (sh.count_fwd > 0 && !overflow) {
      for (uint32_t q = 0; q < sh.n_plen; q++) {
        uint32_t th = thit_get(S, sh, sh.reached_fuz, sh.plen[q]);
        if (th) o.start_pack[q] = S.packidx[th - 1];
      }
    }
    o.nB = sh.nB; o.n_edges = sh.n_edges; o.n_src = sh.n_src; o.n_sink = sh.n_sink;
    o.max_src_level = sh.max_src_level; o.min_sink_level = sh.min_sink_level;
    o.top_level = (uint32_t)sh.top_level;
    o.pack_off = sh.pack_off;
    o.levels_r = (uint32_t)sh.levels_r; o.levels_f = (uint32_t)sh.levels_f;
    o.fill_off = 0; o.fill_len = 0; o.left_fuz = 0; o.rand_calls = 0; o.pad0 = 0;
    if (sh.status & GS_TRACED) {
      o.fill_off = sh.fill_off; o.left_fuz = sh.t_left_fuz;
      o.fill_len = (uint32_t)(sh.plen[0] - (sh.t_lo));
      o.rand_calls = 1u + o.fill_len;
      o.start_pack[0] = o.start_pack[1] = NONE32;
    }
    const unsigned long long t_end = now_cycles();
    // t[0] start, t[1] after init, t[2] before the forward pass, t[3] after it, t[4] after the backward marking
    o.cyc[0] = (uint32_t)((t[1] - t[0]) >> 4); o.cyc[1] = (uint32_t)(sh.cyc_right >> 4);
    o.cyc[2] = (uint32_t)((t[3] - t[2]) >> 4); o.cyc[3] = (uint32_t)((t[4] - t[3]) >> 4);
    o.cyc[4] = (uint32_t)((t_end - t[4]) >> 4); o.cyc[5] = (uint32_t)((t_end - t[0]) >> 4);
    for (int q = 0; q < 5; q++) o.prof[q] = (uint32_t)(sh.prof[q] >> 4);
    for (int q = 5; q < 8; q++) o.prof[q] = (uint32_t)sh.prof[q];
    o.table_lookups = sh.table_lookups; o.node_ops = sh.node_ops;
    *out = o;
  }
  G::sync();
}

// all phases on one group, one after the other (wide tier, host emulation)
template <class G>
G2S_D void process_gap(const FillParams &P, const GapDev &gd, const Slice &S, GShared &sh, GapOut *out) {
  const GapCtx cx = make_ctx(P, gd);
  unsigned long long t[5];
  t[0] = now_cycles();
  gap_init<G>(P, cx, S, sh);
  t[1] = now_cycles();
  gap_right<G>(P, cx, S, sh, sh.nbuf[0]);
  t[2] = now_cycles();
  gap_forward<G, NoWait>(P, gd, cx, S, sh, sh.nbuf[1]);
  t[3] = now_cycles();
  t[4] = gap_backward_pack<G>(P, cx, S, sh, sh.nbuf[1]);
  gap_finish<G>(S, sh, out, t);
}

}  // namespace g2s
