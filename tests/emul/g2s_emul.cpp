// g2s_emul.cpp - TEST INFRASTRUCTURE ONLY.  Compiles the per-gap device algorithm (gap2seq_b200/csrc/g2s_fill.cuh)
// as plain host C++ with a thread group of size 1 (-DG2S_HOST_EMUL) so its sequential logic can be checked
// against the oracle in the CPU test suite (no GPU in the build container).  It shares the algorithm source and
// the host finish stage with libg2s_gpu.so but is never loaded by the product; races and warp-level behaviour
// are only exercised by the `-m gpu` tests on the real kernels.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#include "../../include/g2s.h"
#include "../../gap2seq_b200/csrc/g2s_fill.cuh"
#include "../../gap2seq_b200/csrc/g2s_chain.cuh"
#include "../../gap2seq_b200/csrc/g2s_finish.h"
#include "../../gap2seq_b200/csrc/g2s_scaffold.h"

using namespace g2s;

struct EmulGroup {
  static int rank() { return 0; }
  static int size() { return 1; }
  static void sync() {}
  static bool any(bool p) { return p; }
  static bool any_word(const uint32_t *w, uint32_t bits) { return (*w & bits) != 0; }
  static int min_i32(int v) { return v; }
  static int max_i32(int v) { return v; }
  static uint32_t add_u32(uint32_t v) { return v; }
  static unsigned long long bcast_u64(unsigned long long v) { return v; }
  static constexpr bool kWarp = false, kWarpScout = false;
};

struct EmulGraph {
  std::vector<Slot> slots;
  Table t;
};

static uint32_t host_adj(const Table &t, uint64_t key) {
  uint32_t adj = 0, dummy;
  for (uint32_t nt = 0; nt < 4; nt++) {
    if (table_find(t, canonical(succ_kmer(key, nt, t), t.k).key, &dummy) >= 0) adj |= 1u << nt;
    if (table_find(t, canonical(pred_kmer(key, nt, t), t.k).key, &dummy) >= 0) adj |= 1u << (4 + nt);
  }
  return adj;
}

extern "C" {

void *emul_graph_from_kmers(int k, const uint64_t *kmers, size_t n) {
  EmulGraph *g = new EmulGraph();
  uint64_t cap = 1024;
  while (cap < 2 * n + 2) cap <<= 1;
  g->slots.assign(cap, Slot{KEY_EMPTY, 0, {0, 0}});
  g->t.slots = g->slots.data();
  g->t.cap_mask = cap - 1;
  g->t.k = k;
  g->t.kmask = (1ULL << (2 * k)) - 1;
  for (size_t i = 0; i < n; i++) {
    uint64_t h = mix64(kmers[i]) & g->t.cap_mask;
    while (g->slots[h].key != KEY_EMPTY && g->slots[h].key != kmers[i]) h = (h + 1) & g->t.cap_mask;
    g->slots[h].key = kmers[i];
  }
  for (uint64_t h = 0; h < cap; h++)
    if (g->slots[h].key != KEY_EMPTY) g->slots[h].val = host_adj(g->t, g->slots[h].key);
  std::vector<Slot> hints(cap, Slot{KEY_EMPTY, 0, {0, 0}});
  for (uint64_t h = 0; h < cap; h++)
    if (g->slots[h].key != KEY_EMPTY)
      compute_hints(g->t, g->slots[h].key, (uint32_t)(g->slots[h].val & 0xFF), &hints[h]);
  for (uint64_t h = 0; h < cap; h++)
    if (g->slots[h].key != KEY_EMPTY) { g->slots[h].val = hints[h].val; g->slots[h].h[0] = hints[h].h[0]; g->slots[h].h[1] = hints[h].h[1]; }
  return g;
}

void emul_graph_free(void *g) { delete (EmulGraph *)g; }

void emul_adjacency(void *gp, const uint64_t *kmers, size_t n, uint8_t *out) {
  EmulGraph *g = (EmulGraph *)gp;
  for (size_t i = 0; i < n; i++) { uint32_t a = 0; table_find(g->t, kmers[i], &a); out[i] = (uint8_t)a; }
}

// device pass of a list of gaps, emulated: fills a detached HostBatch exactly like the engine's do_collect
static double g_submit_s = 0, g_analyze_s = 0;
static unsigned long long g_chain_gaps = 0, g_table_gaps = 0;   // gaps finished by the chain tier / by the tiers with tables
static ThreadPool *host_pool() {   // G2S_EMUL_THREADS=1 selects the sequential host stages
  static ThreadPool *p = nullptr;
  static bool init = false;
  if (!init) { init = true; const char *v = getenv("G2S_EMUL_THREADS"); int n = v ? atoi(v) : 4; if (n > 1) p = new ThreadPool(n); }
  return p;
}
static double now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static int emul_submit(EmulGraph *g, int k, const g2s_gap *gaps, size_t n, HostBatch &hb, int cap0) {
  const double t_submit0 = now_s();
  hb.k = k;
  hb.gaps.assign(gaps, gaps + n);
  hb.outs.assign(n, GapOut{});
  hb.tier.assign(n, 0);
  hb.pack_of.assign(n, 0);
  hb.own.assign(n, std::vector<PackRec>());
  hb.pack_base.assign(n, nullptr);
  hb.fill_base.assign(n, nullptr);
  hb.own_fill.assign(n, std::vector<char>());   // fills of the gaps the (emulated) device traced itself
  const bool device_trace = getenv("G2S_NO_DEVICE_TRACE") == nullptr;
  for (size_t i = 0; i < n; i++) {
    const g2s_gap &gg = gaps[i];
    std::vector<char> flanks((size_t)3 * k + gg.left_max_fuz + 2 * gg.right_max_fuz + 1);
    memcpy(flanks.data(), gg.kmer_left, (size_t)k + gg.left_max_fuz);
    memcpy(flanks.data() + k + gg.left_max_fuz, gg.kmer_right, (size_t)k + gg.right_max_fuz);
    GapDev gd{};
    gd.rstart_off = (uint32_t)(k + gg.left_max_fuz);
    if (gg.kmer_right_starts) {
      gd.rstart_off = (uint32_t)(2 * k + gg.left_max_fuz + gg.right_max_fuz);
      memcpy(flanks.data() + gd.rstart_off, gg.kmer_right_starts, (size_t)k + gg.right_max_fuz);
    }
    gd.flank_off = 0; gd.gap_len = gg.gap_len; gd.gap_err = gg.gap_err; gd.lmf = gg.left_max_fuz; gd.rmf = gg.right_max_fuz;
    gd.flags = (uint32_t)gg.flags; gd.id = (uint32_t)i;
    const int dmax = gg.gap_len + gg.gap_err + gg.left_max_fuz + gg.right_max_fuz;
    uint64_t cap = cap0 > 0 ? (uint64_t)cap0 : (uint64_t)4 * (dmax + 2) + 1024;
    GapOut out{};
    std::vector<PackRec> pack;
    int tier = 0;
    // the chain tier first (g2s_chain.cuh); G2S_NO_CHAIN=1 leaves every gap to the tiers with tables
    static const bool chain_on = getenv("G2S_NO_CHAIN") == nullptr;
    if (chain_on && device_trace && cap0 == 0) {
      const int right_half = gg.right_max_fuz + (gg.gap_len + gg.gap_err + 1) / 2;
      ChainParams CP{};
      CP.table = g->t; CP.flanks = flanks.data(); CP.gaps = &gd; CP.n_gaps = 1; CP.outs = &out;
      CP.f_words = chain_words(k, dmax); CP.r_words = chain_words(k, right_half);
      std::vector<uint64_t> strs((size_t)CP.f_words + CP.r_words, 0xDEADBEEFDEADBEEFULL);
      ChainWalk walks[2];
      std::vector<char> fills((size_t)dmax + 24);
      unsigned long long fcursor = 0;
      CP.strs = strs.data(); CP.walks = walks; CP.fills = fills.data(); CP.fill_cap = fills.size(); CP.fill_cursor = &fcursor;
      chain_walk(g->t, kmer_from_ascii(flanks.data() + gd.flank_off, k), gg.gap_len / 2 + gg.gap_err / 2 + gg.left_max_fuz, strs.data(), &walks[0], false);
      chain_walk(g->t, revcomp(kmer_from_ascii(flanks.data() + gd.rstart_off + gd.rmf, k), k), right_half, strs.data() + CP.f_words, &walks[1], true);
      uint64_t tk[4 * (MAX_FUZ + 1)];
      chain_finish<EmulGroup>(CP, gd, strs.data(), strs.data() + CP.f_words, walks[0], walks[1], tk, &out);
      if (out.status & GS_INTERNAL) return -100;
      if (!(out.status & GS_NOT_CHAIN)) {
        g_chain_gaps++;
        if (out.status & GS_TRACED) { hb.own_fill[i].assign(fills.begin() + out.fill_off, fills.begin() + out.fill_off + out.fill_len + 1); out.fill_off = 0; }
        hb.outs[i] = out; hb.tier[i] = -1; hb.pack_of[i] = (int)i;
        hb.pack_base[i] = nullptr;
        if (out.status & GS_TRACED) hb.fill_base[i] = hb.own_fill[i].data();
        hb.gaps[i].kmer_left = hb.gaps[i].kmer_right = nullptr;
        continue;
      }
      out = GapOut{};
    }
    g_table_gaps++;
    for (;; tier++, cap *= 16) {
      uint32_t hlog = tier == 0 ? ceil_log2(3 * ((uint64_t)dmax + 64)) : (ceil_log2(cap / 4) > 16 ? ceil_log2(cap / 4) : 16);
      SliceLayout L = make_layout(cap, hlog, (uint32_t)dmax + 1, (uint32_t)std::max(gg.left_max_fuz, gg.right_max_fuz) + 1);
      std::vector<unsigned char> scratch(L.bytes);
      pack.assign(cap, PackRec{});
      unsigned long long cursor = 0, fcursor = 0;
      std::vector<char> fills((size_t)dmax + 8);
      uint32_t order = 0, qh = 0;
      FillParams P{};
      P.table = g->t; P.flanks = flanks.data(); P.gaps = &gd; P.order = &order; P.n_gaps = 1; P.outs = &out;
      P.scratch = scratch.data(); P.lay = L; P.pack = pack.data(); P.pack_cap = cap; P.pack_cursor = &cursor; P.queue_head = &qh;
      if (device_trace) { P.fills = fills.data(); P.fill_cap = fills.size(); P.fill_cursor = &fcursor; }
      GShared sh{};
      Slice S = make_slice(scratch.data(), L);
      process_gap<EmulGroup>(P, gd, S, sh, &out);
      if (out.status & GS_INTERNAL) return -100;
      if (!(out.status & (GS_OVERFLOW | GS_OUT_FULL))) {
        if (out.status & GS_TRACED) { hb.own_fill[i].assign(fills.begin() + out.fill_off, fills.begin() + out.fill_off + out.fill_len + 1); out.fill_off = 0; }
        break;
      }
      if (cap > (1ULL << 28)) return -101;
    }
    pack.resize(out.npack + out.pack_off);
    hb.outs[i] = out;
    hb.tier[i] = tier;
    hb.pack_of[i] = (int)i;
    hb.own[i].swap(pack);
    hb.pack_base[i] = hb.own[i].data();
    if (out.status & GS_TRACED) hb.fill_base[i] = hb.own_fill[i].data();
    hb.gaps[i].kmer_left = hb.gaps[i].kmer_right = nullptr;
  }
  const double t_an0 = now_s();
  hb.analyze(host_pool());
  g_analyze_s += now_s() - t_an0;
  g_submit_s += now_s() - t_submit0;
  return 0;
}
extern "C" void emul_tier_counts(unsigned long long *chain, unsigned long long *table, int reset) {
  *chain = g_chain_gaps; *table = g_table_gaps;
  if (reset) g_chain_gaps = g_table_gaps = 0;
}
extern "C" void emul_times(double *submit_s, double *analyze_s, int reset) {
  *submit_s = g_submit_s; *analyze_s = g_analyze_s;
  if (reset) g_submit_s = g_analyze_s = 0;
}

// same contract as g2s_fill_batch; cap0 > 0 forces the first tier's cell capacity (to exercise tier escalation)
int emul_fill_batch(void *gp, int k, const g2s_gap *gaps, size_t n, g2s_result *results, char *const *fill_out, int cap0) {
  HostBatch hb;
  int rc = emul_submit((EmulGraph *)gp, k, gaps, n, hb, cap0);
  if (rc) return rc;
  std::vector<TraceJob> jobs(n);
  for (size_t i = 0; i < n; i++) jobs[i] = TraceJob{&hb, i};
  std::vector<g2s_result> res;
  RawBuf<char> arena_buf;
  std::vector<size_t> off;
  trace_many(jobs, res, arena_buf, off, host_pool(), nullptr);
  const char *arena = arena_buf.p;
  for (size_t i = 0; i < n; i++) {
    results[i] = res[i];
    if (fill_out && fill_out[i]) memcpy(fill_out[i], &arena[off[i]], (size_t)res[i].fill_len + 1);
  }
  return 0;
}

// same contract as g2s_fill_records
int emul_fill_records(void *gp, const char *const *seqs, const char *const *comments, size_t n_records,
                      const g2s_run_options *opt, char **filled_out, char **log_out, int32_t *gapcount, int32_t *filledcount,
                      g2s_result **gap_results, size_t *n_gap_results) {
  std::vector<StrRef> s(n_records), c(n_records);
  for (size_t i = 0; i < n_records; i++) { s[i] = StrRef(seqs[i]); c[i] = StrRef(comments[i]); }
  RunOptions ro;
  ro.k = opt->k; ro.d_err = opt->dist_error; ro.max_fuz = opt->fuz;
  ro.skip_confident = opt->skip_confident != 0; ro.unique_paths = opt->unique_paths != 0; ro.all_paths = opt->all_paths != 0;
  ScaffoldRun run(s, c, ro);
  run.pool = host_pool();
  SubmitFn submit = [&](const std::vector<g2s_gap> &gaps, HostBatch &hb) -> int {
    return emul_submit((EmulGraph *)gp, opt->k, gaps.data(), gaps.size(), hb, 0);
  };
  std::vector<std::string> filled;
  std::string log;
  int gc = 0, fc = 0;
  std::vector<g2s_result> details;
  int rc = run.run(submit, filled, log, gc, fc, &details);
  if (rc) return rc;
  if (getenv("G2S_HOST_PROF")) fprintf(stderr, "emul_fill_records: phase1 %.3f ms (incl. emulated device pass), trace+splice+log %.3f ms\n", run.phase1_ms, run.phase2_ms);
  for (size_t i = 0; i < n_records; i++) filled_out[i] = strdup(filled[i].c_str());
  *log_out = strdup(log.c_str());
  *gapcount = gc; *filledcount = fc;
  if (gap_results) {
    *n_gap_results = details.size();
    *gap_results = (g2s_result *)malloc((details.size() + 1) * sizeof(g2s_result));
    memcpy(*gap_results, details.data(), details.size() * sizeof(g2s_result));
  }
  return 0;
}

void emul_free(void *p) { free(p); }

}  // extern "C"

// debugging aid: histogram of skip-hint lengths (HINT_MAX + 1 bins per orientation) over all nodes
extern "C" void emul_hint_hist(void *gp, uint64_t *hist82) {
  EmulGraph *g = (EmulGraph *)gp;
  for (int i = 0; i < 2 * (HINT_MAX + 1); i++) hist82[i] = 0;
  for (const Slot &s : g->slots) {
    if (s.key == KEY_EMPTY) continue;
    hist82[hint_of(s, 0).len]++;
    hist82[HINT_MAX + 1 + hint_of(s, 1).len]++;
  }
}

extern "C" int emul_rand_skip_enabled() { return g2s::RandSkipper::enabled() ? 1 : 0; }
extern "C" void emul_rand_skip(size_t n) { g2s::RandSkipper::skip(n); }

#ifdef G2S_EMUL_STATS
// development aid (tools/emul_stats.py): probe / level-width counters of the emulated pass
extern "C" void emul_stats_get(unsigned long long *out, int reset) {
  memcpy(out, &g2s::emul_stats(), sizeof(g2s::EmulStats));
  if (reset) memset(&g2s::emul_stats(), 0, sizeof(g2s::EmulStats));
}
#endif

extern "C" int emul_rand_ahead_enabled() { return g2s::RandAhead::enabled() ? 1 : 0; }
extern "C" void emul_trace_stats(unsigned long long *out4) {
  g2s::TraceStats &t = g2s::trace_stats();
  out4[0] = t.calls; out4[1] = t.parallel_calls; out4[2] = t.rounds; out4[3] = t.retraced;
}

