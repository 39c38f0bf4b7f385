// g2s_finish.h - the host stages that follow the device pass of one gap (host-only code, part of the product):
//   analyze_general  SCC contraction + branch counter of src/Gap2Seq.cpp:1314-1434 for path subgraphs that are
//                    not a plain level DAG (the device classifies the DAG case itself, g2s_fill.cuh)
//   finish_one       result assembly + the random traceback of src/Gap2Seq.cpp:1437-1517 over the device-built
//                    predecessor trellis; the only consumer of libc rand()
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/g2s.h"
#include "g2s_fill.cuh"
#include "g2s_pool.h"

namespace g2s {

// Evicting a byte range from the CPU caches.  The pack arenas arrive by DMA in pinned staging buffers that the engine
// reuses; once the traceback threads have read a buffer its lines sit in several cores' caches and the next DMA into
// it pays a snoop per line (measured on the B200 host: 5x slower D2H with 16 reader threads).  Flushing the lines
// when a batch is done - in parallel, they are clean - keeps the next D2H at PCIe speed.
#if defined(__x86_64__)
__attribute__((target("clflushopt"))) inline void flush_lines_opt(const char *p, const char *e) {
  for (; p < e; p += 64) __builtin_ia32_clflushopt((void *)p);
  __builtin_ia32_sfence();
}
inline void flush_lines_plain(const char *p, const char *e) {
  for (; p < e; p += 64) __builtin_ia32_clflush((const void *)p);
}
inline bool cpu_has_clflushopt() {
  static const bool has = [] {
    unsigned a = 7, b = 0, c = 0, d = 0;
    __asm__ volatile("cpuid" : "+a"(a), "=b"(b), "+c"(c), "=d"(d));
    (void)d;
    return ((b >> 23) & 1u) != 0;
  }();
  return has;
}
inline void flush_cache_range(const void *ptr, size_t bytes) {
  const char *p = (const char *)((uintptr_t)ptr & ~(uintptr_t)63), *e = (const char *)ptr + bytes;
  if (cpu_has_clflushopt()) flush_lines_opt(p, e); else flush_lines_plain(p, e);
}
#else
inline void flush_cache_range(const void *, size_t) {}
#endif

struct GeneralResult {
  g2s_subgraph_stats st{};
  std::vector<uint8_t> safe;   // per pack record
};

// SCC contraction + branch counter for a subgraph that is not a level DAG (src/Gap2Seq.cpp:1314-1434).
// Vertices are the distinct nodes of the B records plus sink (0) and source (1); edges are de-duplicated
// exactly like the reference's `if (!edge(u,v).second) add_edge` (Gap2Seq.cpp:1222-1297).
inline void analyze_general(const PackRec *pk, uint32_t npack, GeneralResult &R) {
  // dense vertex numbers in first-seen order (every result below is invariant under renumbering):
  // node-table entry -> vertex through a direct map, then once per record so that edges never search
  // the B records (vertices of the subgraph) are a fraction of the pack - the rest is the traceback trellis - so they
  // are listed once and every later pass walks the list instead of streaming all 32-byte records again
  std::vector<uint32_t> bl;
  bl.reserve(npack / 4 + 16);
  uint32_t maxnode = 0;
  for (uint32_t i = 0; i < npack; i++)
    if (pk[i].meta & PM_B) { bl.push_back(i); if (pk[i].node > maxnode) maxnode = pk[i].node; }
  std::vector<uint32_t> vmap((size_t)maxnode + 1, NONE32), vof(npack, NONE32);
  uint32_t V = 2;   // 0 = sink, 1 = source
  for (uint32_t i : bl) {
    uint32_t &m = vmap[pk[i].node];
    if (m == NONE32) m = V++;
    vof[i] = m;
  }
  // out-edges bucketed by source (counting sort), duplicates - the same node pair on several levels - removed per
  // bucket: the reference's `if (!edge(u,v).second) add_edge`
  std::vector<uint32_t> ostart((size_t)V + 1, 0);
  for (uint32_t i : bl) {
    if (pk[i].meta & PM_START) ostart[1 + 1]++;
    else for (int q = 0; q < 4; q++) if (pk[i].pred[q] != NONE32) ostart[vof[pk[i].pred[q]] + 1]++;
    if (pk[i].meta & PM_SEED) ostart[vof[i] + 1]++;
  }
  for (uint32_t v = 0; v < V; v++) ostart[v + 1] += ostart[v];
  std::vector<uint32_t> otgt(ostart[V]), fillp(ostart.begin(), ostart.end() - 1);
  for (uint32_t i : bl) {
    const uint32_t v = vof[i];
    if (pk[i].meta & PM_START) otgt[fillp[1]++] = v;
    else for (int q = 0; q < 4; q++) if (pk[i].pred[q] != NONE32) otgt[fillp[vof[pk[i].pred[q]]]++] = v;
    if (pk[i].meta & PM_SEED) otgt[fillp[v]++] = 0u;
  }
  {
    uint32_t w = 0;
    for (uint32_t v = 0; v < V; v++) {
      const uint32_t b0 = ostart[v], e0 = ostart[v + 1];
      ostart[v] = w;
      if (e0 - b0 > 1) std::sort(otgt.begin() + b0, otgt.begin() + e0);
      for (uint32_t j = b0; j < e0; j++) if (j == b0 || otgt[j] != otgt[j - 1]) otgt[w++] = otgt[j];
    }
    ostart[V] = w;
    otgt.resize(w);
  }
  const size_t E = otgt.size();

  // iterative Tarjan
  std::vector<uint32_t> index(V, NONE32), low(V, 0), comp(V, NONE32), pos(V, 0), stack, call;
  std::vector<uint8_t> on(V, 0);
  uint32_t counter = 0, ncomp = 0;
  for (uint32_t r = 0; r < V; r++) {
    if (index[r] != NONE32) continue;
    call.push_back(r); index[r] = low[r] = counter++; stack.push_back(r); on[r] = 1; pos[r] = ostart[r];
    while (!call.empty()) {
      uint32_t v = call.back();
      if (pos[v] < ostart[v + 1]) {
        uint32_t w = otgt[pos[v]++];
        if (index[w] == NONE32) { index[w] = low[w] = counter++; stack.push_back(w); on[w] = 1; pos[w] = ostart[w]; call.push_back(w); }
        else if (on[w]) low[v] = std::min(low[v], index[w]);
      } else {
        call.pop_back();
        if (!call.empty()) low[call.back()] = std::min(low[call.back()], low[v]);
        if (low[v] == index[v]) {
          for (;;) { uint32_t w = stack.back(); stack.pop_back(); on[w] = 0; comp[w] = ncomp; if (w == v) break; }
          ncomp++;
        }
      }
    }
  }
  std::vector<uint32_t> csize(ncomp, 0);
  for (uint32_t v = 0; v < V; v++) csize[comp[v]]++;
  uint64_t nontriv = 0, size_nontriv = 0, self_loops_trivial = 0, cross = 0;
  for (uint32_t c = 0; c < ncomp; c++) if (csize[c] > 1) { nontriv++; size_nontriv += csize[c]; }
  // condensed multigraph: every cross-component edge exactly once (Gap2Seq.cpp:1346-1370)
  std::vector<uint32_t> indeg(ncomp, 0), outdeg(ncomp, 0);
  for (uint32_t u = 0; u < V; u++) {
    for (uint32_t j = ostart[u]; j < ostart[u + 1]; j++) {
      const uint32_t v = otgt[j];
      if (comp[u] == comp[v]) { if (u == v && csize[comp[u]] == 1) self_loops_trivial++; continue; }
      cross++; outdeg[comp[u]]++; indeg[comp[v]]++;
    }
  }
  R.st.vertices = V;
  R.st.edges = E - self_loops_trivial;
  R.st.nontrivial_components = nontriv;
  R.st.size_nontrivial_components = size_nontriv;
  R.st.vertices_final = V + nontriv - size_nontriv;
  R.st.edges_final = cross;
  // branch counter in a topological order of the condensation (order invariant).  Tarjan numbers
  // components in reverse topological order, so descending component id is a valid order.
  std::vector<int> branch(ncomp, 0);
  long long bc = 1;
  for (uint32_t c = ncomp; c-- > 0;) {
    if (indeg[c] == 0 && outdeg[c] == 0) continue;
    if (indeg[c] > 1) bc -= (long long)indeg[c] - 1;
    branch[c] = (int)bc;
    if (outdeg[c] > 1) bc += (long long)outdeg[c] - 1;
  }
  auto vertex_safe = [&](uint32_t v) { return csize[comp[v]] == 1 && branch[comp[v]] == 1; };
  const bool sink_safe = vertex_safe(0);
  R.safe.assign(npack, 0);
  for (uint32_t i = 0; i < npack; i++) {
    const uint32_t v = pk[i].node <= maxnode ? vmap[pk[i].node] : NONE32;
    R.safe[i] = v != NONE32 ? vertex_safe(v) : sink_safe;   // node2boost[] default-inserts the sink
  }
}


// Consuming libc rand() values whose result is not needed.  The reference calls rand() once per traceback step even
// when the step has a single predecessor (src/Gap2Seq.cpp:1513), so >99 % of the calls on typical gaps only advance the
// generator - at ~20 ns per locked glibc call that is most of the host time.  glibc's default generator is the
// additive feedback r[i] = r[i-3] + r[i-31] over a 31-word table that initstate()/setstate() expose: setstate()
// stores the rear index in the word before the table and reloads it from there.  skip(n) switches libc to a dummy
// state, advances the application's table n steps in place (one add per step, no lock), and switches back, so the
// next rand() returns exactly what it would have returned after n discarded calls.  A self test (on a private
// state, the application's stream is restored) decides once whether this libc behaves that way; if not, skip(n) just
// calls rand() n times.
class RandAhead;
class RandSkipper {
 public:
  static void skip(size_t n) {
    if (n == 0) return;
    if (n < 4 || !enabled()) { while (n--) rand(); return; }
    if (!fast_skip(n)) while (n--) rand();
  }
  static bool enabled() {
    static std::once_flag once;
    static bool ok = false;
    std::call_once(once, [] { ok = self_test(); });
    return ok;
  }

 private:
  static char *dummy() {
    static int32_t buf[34];
    static bool init = false;
    if (!init) { char *app = initstate(1u, (char *)buf, 128); setstate(app); init = true; }
    return (char *)buf;
  }
  static bool fast_skip(size_t n) {
    int32_t *info = (int32_t *)setstate(dummy());   // libc writes its rear index into info[0] and lets go of the table
    if (!info) return false;
    const int type = info[0] % 5, rear = info[0] / 5;
    if (type != 3 || rear < 0 || rear >= 31) { setstate((char *)info); return false; }
    uint32_t *st = (uint32_t *)(info + 1);
    int f = (rear + 3) % 31, r = rear;
    for (size_t i = 0; i < n; i++) {
      st[f] += st[r];
      if (++f == 31) f = 0;
      if (++r == 31) r = 0;
    }
    info[0] = r * 5 + 3;
    setstate((char *)info);
    return true;
  }
  friend class RandAhead;
  static bool self_test() {
    static int32_t bufa[34];
    char *app = initstate(777u, (char *)bufa, 128);
    if (!app) return false;
    int ref[160];
    for (int i = 0; i < 160; i++) ref[i] = rand();
    initstate(777u, (char *)bufa, 128);
    bool ok = fast_skip(133);
    for (int i = 133; i < 160 && ok; i++) ok = rand() == ref[i];
    if (ok) { ok = fast_skip(5); for (int i = 0; i < 3 && ok; i++) rand(); }
    setstate(app);   // the application's stream continues exactly where it was
    return ok;
  }
};

// The sources of random numbers a traceback can draw from.  LibcRand is the application's generator itself;
// RandCursor reads a window of that same stream that RandAhead computed beforehand, which is what lets tracebacks of
// different gaps run on different threads: gap j starts at the stream position where gap j-1 stopped.
struct LibcRand {
  int next() { return rand(); }
  void skip(size_t n) { RandSkipper::skip(n); }
};
// grow-only uninitialised storage that survives from call to call: fresh multi-megabyte allocations cost a page fault
// per 4 KB on first touch, which on the measured host was several times the work done in them
template <class T>
struct RawBuf {
  T *p = nullptr;
  size_t cap = 0;
  RawBuf() {}
  RawBuf(const RawBuf &) = delete;
  RawBuf &operator=(const RawBuf &) = delete;
  ~RawBuf() { free(p); }
  T *ensure(size_t n, size_t keep = 0) {   // room for n elements; the first `keep` survive a reallocation
    if (n > cap) {
      const size_t ncap = n + n / 2 + 64;
      T *q = (T *)malloc(ncap * sizeof(T));
      if (p && keep) memcpy(q, p, keep * sizeof(T));
      free(p);
      p = q; cap = ncap;
    }
    return p;
  }
};

class RandAhead;
struct RandCursor {
  const uint32_t *s;   // generator words of the stream (RandAhead); rand() returns word >> 1
  size_t pos;
  RandAhead *owner;    // materialises the block of the stream a value is read from (skipped stretches never exist)
  inline int next();
  void skip(size_t n) { pos += n; }
};

// The next values of libc rand(), computed from a private copy of the generator's table without consuming them.
// glibc's default generator is the additive feedback s[i] = s[i-3] + s[i-31] (mod 2^32), rand() = s[i] >> 1, kept in a
// circular table of 31 words; laid out as a plain array in stream order the recurrence is one add per value with no
// index arithmetic.
// The reference draws one value per traceback step even where the step has no choice (src/Gap2Seq.cpp:1513), so a batch
// of 10 000 gaps spans tens of millions of values of which almost none is ever looked at.  The stream is therefore laid
// out in BLOCKS of 65 536 values that are computed only when a traceback reads from them: extend(n) merely computes the
// 31-word generator state in front of every block - a jump by a whole block is a product with x^65536 mod the
// generator's characteristic polynomial x^31 - x^28 - 1 over Z/2^32, 961 multiply-adds - and need(lo, hi) fills the
// blocks that overlap [lo, hi), from any thread.  commit(used) writes the 31 words in front of position `used` back
// into libc's table and moves its rear index, so the application's stream continues exactly as after `used` rand()
// calls.  Only valid where RandSkipper's self test recognised that generator; enabled() checks the whole mechanism
// (values across block boundaries, commit) once more against rand() itself.
class RandAhead {
 public:
  static const size_t BLOCK = (size_t)1 << 16;
  static bool enabled() {
    static std::once_flag once;
    static bool good = false;
    std::call_once(once, [] {
      if (!RandSkipper::enabled()) return;
      static int32_t bufa[34];
      char *app = initstate(4242u, (char *)bufa, 128);
      if (!app) return;
      bool ok = true;
      for (size_t n : {(size_t)5, (size_t)31, (size_t)8269, 3 * BLOCK + 77}) {
        std::vector<int> ref(n + 8);
        initstate(4242u, (char *)bufa, 128);
        for (size_t i = 0; i < n + 8; i++) ref[i] = rand();
        initstate(4242u, (char *)bufa, 128);
        RandAhead a;
        ok = ok && a.begin_unchecked();
        if (ok) { a.extend(3); a.extend(n + 8); }
        if (ok && n > BLOCK) {   // blocks out of order, a block in the middle never: values and commit must not depend on it
          a.need(2 * BLOCK + 5, n + 8);
          for (size_t i = 2 * BLOCK + 5; i < n + 8 && ok; i++) ok = (int)(a.data()[i] >> 1) == ref[i];
          a.need(0, 100);
          for (size_t i = 0; i < 100 && ok; i++) ok = (int)(a.data()[i] >> 1) == ref[i];
        } else if (ok) {
          a.need(0, n + 8);
          for (size_t i = 0; i < n + 8 && ok; i++) ok = (int)(a.data()[i] >> 1) == ref[i];
        }
        if (ok) ok = a.commit(n);
        for (size_t i = n; i < n + 8 && ok; i++) ok = rand() == ref[i];
      }
      setstate(app);
      good = ok;
    });
    return good;
  }
  bool begin() { return enabled() && begin_unchecked(); }
  bool begin_unchecked() {
    n_ = 0; ok_ = false;
    if (!RandSkipper::enabled()) return false;
    int32_t *info = (int32_t *)setstate(RandSkipper::dummy());
    if (!info) return false;
    const int type = info[0] % 5, rear = info[0] / 5;
    if (type == 3 && rear >= 0 && rear < 31) {
      const uint32_t *st = (const uint32_t *)(info + 1);
      r0_ = rear; f0_ = (rear + 3) % 31;
      states_.assign(1, State());
      for (int j = -31; j < 0; j++) states_[0].w[31 + j] = st[((f0_ + j) % 31 + 31) % 31];   // word of step j sits at (f0 + j) mod 31
      nblk_ = 0;
      ok_ = true;
    }
    setstate((char *)info);
    return ok_;
  }
  bool ok() const { return ok_; }
  size_t size() const { return n_; }
  const uint32_t *data() const { return buf_.p; }
  // make positions [0, n) addressable: room in the flat array, the generator state in front of each of their blocks.
  // Not to be called while tracebacks read (the array may move).
  void extend(size_t n) {
    if (!ok_ || n <= n_) return;
    const size_t nb = (n + BLOCK - 1) / BLOCK;
    bool fresh = false;
    if (nb * BLOCK > buf_.cap) { buf_.ensure(nb * BLOCK + BLOCK); fresh = true; }   // a new array: nothing materialised is
    if (flags_.size() < nb) {                                                       // carried over (it is cheap to redo)
      std::vector<std::atomic<int>> f(nb + nb / 2 + 4);
      flags_.swap(f);
      fresh = true;
    }
    for (size_t b = fresh ? 0 : nblk_; b < (fresh ? flags_.size() : nb); b++) flags_[b].store(0, std::memory_order_relaxed);
    const uint32_t *q = jump_poly();
    while (states_.size() < nb) {          // state in front of block b + 1 = x^BLOCK applied to the state in front of block b
      const State &a = states_.back();
      uint32_t w[61];
      for (int i = 0; i < 31; i++) w[i] = a.w[i];
      for (int i = 31; i < 61; i++) w[i] = w[i - 31] + w[i - 3];
      State nx;
      for (int j = 0; j < 31; j++) { uint32_t acc = 0; for (int m = 0; m < 31; m++) acc += q[m] * w[j + m]; nx.w[j] = acc; }
      states_.push_back(nx);
    }
    nblk_ = nb;
    n_ = n;
  }
  // the words of positions [lo, hi) are about to be read
  void need(size_t lo, size_t hi) {
    if (!ok_ || hi <= lo) return;
    if (hi > n_) hi = n_;
    for (size_t b = lo / BLOCK; b * BLOCK < hi; b++) fill_block(b);
  }
  void touch(size_t pos) { if (flag(pos / BLOCK).load(std::memory_order_acquire) != 2) fill_block(pos / BLOCK); }
  // advance libc's generator by `used` <= size() values
  bool commit(size_t used) {
    if (!ok_) return false;
    if (used == 0) return true;
    if (used > n_) return false;
    need(used > 31 ? used - 31 : 0, used);
    int32_t *info = (int32_t *)setstate(RandSkipper::dummy());
    if (!info) return false;
    uint32_t *st = (uint32_t *)(info + 1);
    const uint32_t *sw = data();
    for (size_t j = used > 31 ? used - 31 : 0; j < used; j++) st[(f0_ + j) % 31] = sw[j];
    info[0] = (int32_t)((r0_ + used) % 31) * 5 + 3;
    setstate((char *)info);
    ok_ = false;
    return true;
  }

 private:
  struct State { uint32_t w[31]; };        // the 31 words in front of a block, oldest first
  std::atomic<int> &flag(size_t b) { return flags_[b]; }
  // x^BLOCK mod (x^31 - x^28 - 1) over Z/2^32, by repeated squaring
  static const uint32_t *jump_poly() {
    static uint32_t q[31];
    static std::once_flag once;
    std::call_once(once, [] {
      uint32_t r[31] = {0};
      r[1] = 1;
      for (size_t e = 1; e < BLOCK; e <<= 1) {
        uint32_t t[61] = {0};
        for (int i = 0; i < 31; i++) for (int j = 0; j < 31; j++) t[i + j] += r[i] * r[j];
        for (int d = 60; d >= 31; d--) { const uint32_t c = t[d]; t[d - 3] += c; t[d - 31] += c; }   // x^31 = x^28 + 1
        for (int i = 0; i < 31; i++) r[i] = t[i];
      }
      for (int i = 0; i < 31; i++) q[i] = r[i];
    });
    return q;
  }
  void fill_block(size_t b) {
    std::atomic<int> &f = flag(b);
    int st = f.load(std::memory_order_acquire);
    if (st == 2) return;
    int expect = 0;
    if (!f.compare_exchange_strong(expect, 1, std::memory_order_acq_rel)) {   // another thread is at it (or done)
      while (f.load(std::memory_order_acquire) != 2) std::this_thread::yield();
      return;
    }
    uint32_t *w = buf_.p + b * BLOCK;
    const State &h = states_[b];
    // the first 31 words read the stored history, then the recurrence runs inside the block: the three most recent words
    // stay in registers, so that the loop is one load, one add and one store per word
    uint32_t tmp[62];
    for (int i = 0; i < 31; i++) tmp[i] = h.w[i];
    for (int i = 31; i < 62; i++) tmp[i] = tmp[i - 31] + tmp[i - 3];
    for (int i = 0; i < 31; i++) w[i] = tmp[31 + i];
    size_t i = 31;
    uint32_t a0 = w[i - 3], a1 = w[i - 2], a2 = w[i - 1];
    for (; i + 3 <= BLOCK; i += 3) {
      a0 += w[i - 31]; a1 += w[i - 30]; a2 += w[i - 29];
      w[i] = a0; w[i + 1] = a1; w[i + 2] = a2;
    }
    for (; i < BLOCK; i++) w[i] = w[i - 31] + w[i - 3];
    f.store(2, std::memory_order_release);
  }

  RawBuf<uint32_t> buf_;        // the stream, block by block where materialised; kept from call to call
  std::vector<State> states_;
  std::vector<std::atomic<int>> flags_;   // per block: 0 not computed, 1 being computed, 2 there
  size_t n_ = 0, nblk_ = 0;
  int f0_ = 0, r0_ = 0;
  bool ok_ = false;
};
inline int RandCursor::next() { if (owner) owner->touch(pos); return (int)(s[pos++] >> 1); }

// Assemble the g2s_result of one gap and run its traceback.  `fill_out` (may be null) receives what the
// reference leaves at &fill[left_max_fuz - left_fuz].
template <class RS>
inline void finish_one(const g2s_gap &g, const GapOut &o, const PackRec *pk, int k, int tier, const GeneralResult &gen,
                       g2s_result &r, char *fill_out, std::vector<char> &buf, RS &rs, const char *dev_fill = nullptr) {
  static const char NT[4] = {'A', 'C', 'T', 'G'};
  memset(&r, 0, sizeof(r));
  const bool skip_confident = (g.flags & G2S_GAP_SKIP_CONFIDENT) != 0, all_paths = (g.flags & G2S_GAP_ALL_PATHS) != 0;
  r.cells_right = o.cells_right; r.cells_left = o.cells_left; r.nodes = o.nodes;
  r.forward_count = (int32_t)o.count_fwd;
  r.n_path_lengths = (int32_t)o.n_plen; r.path_lengths[0] = o.plen[0]; r.path_lengths[1] = o.plen[1];
  r.reached_fuz = o.reached_fuz;
  if (o.status & GS_ORDER_DEPENDENT) r.status |= G2S_ST_ORDER_DEPENDENT;
  if (tier > 0) r.status |= G2S_ST_BIG_TIER;
  if (o.count_fwd == 0 || o.n_plen == 0) return;
  const bool general = (o.status & GS_GENERAL) != 0 && !skip_confident;
  int count = (int)o.count_fwd;
  r.right_fuz = o.reached_fuz;
  if (!skip_confident) {
    if (general) {
      r.substats = gen.st;
      r.status |= G2S_ST_GENERAL_SUBGRAPH;
    } else {   // level DAG: no contraction happens, the device counted vertices and edges
      r.substats.vertices = (uint64_t)o.nB + 2;
      r.substats.edges = (uint64_t)o.n_edges + o.n_src + o.n_sink;
      r.substats.vertices_final = r.substats.vertices;
      r.substats.edges_final = r.substats.edges;
    }
    r.substats_valid = 1;
    if (all_paths) count = (int)o.count_all;   // the Phase B recount replaces the first-found count (Gap2Seq.cpp:1192)
  }
  if (o.status & GS_TRACED) {   // no choice anywhere: the device wrote the fill, only the rand() stream moves on
    if (!dev_fill) { r.status |= G2S_ST_UNABLE_TO_BACKTRACE; r.count = 0; return; }   // cannot happen: the arena travels with the batch
    r.status |= G2S_ST_DEVICE_TRACED;
    r.rand_calls = (int32_t)o.rand_calls;
    rs.skip(o.rand_calls);
    r.left_fuz = o.left_fuz;
    r.fill_len = (int32_t)o.fill_len;
    if (fill_out) { memcpy(fill_out, dev_fill, (size_t)o.fill_len); fill_out[o.fill_len] = 0; }
    r.count = count;
    return;
  }
  r.rand_calls = 1;
  size_t pending = 0;   // rand() values whose result cannot matter (a single choice): consumed in bulk
  uint32_t q = 0;
  if (o.n_plen > 1) q = (uint32_t)(rs.next() % (int)o.n_plen);   // Gap2Seq.cpp:1440
  else pending++;
  int d = o.plen[q];
  const int d0 = d;
  int last_solid = d;
  uint32_t cur = o.start_pack[q];
  buf.assign((size_t)d0 + 1, 0);
  int left_fuz = 0;
  bool broken = (cur == NONE32);
  while (d >= 0 && !broken) {
    const PackRec &rec = pk[cur];
    // predecessors usually follow in the arena (it is filled top level first).  Non-temporal: the records are read
    // once, and lines left in the cache hierarchy slow down the next DMA into the staging buffer
    __builtin_prefetch(&pk[cur + 24], 0, 0);
    if (d <= g.left_max_fuz && (rec.meta & PM_START)) { left_fuz = g.left_max_fuz - d; break; }
    if (d > 0) {
      const bool safe = skip_confident || (general ? gen.safe[cur] != 0 : (rec.meta & PM_SAFE) != 0);
      if (safe) last_solid = d;
      const char c = NT[(rec.meta >> PM_BASE_SHIFT) & 3];
      buf[d - 1] = (d > last_solid - k) ? c : (char)(c | 0x20);
      uint32_t back[4]; int nb = 0;
      for (int t = 0; t < 4; t++) if (rec.pred[t] != NONE32) back[nb++] = rec.pred[t];
      if (nb == 0) { broken = true; break; }
      r.rand_calls++;            // one rand() per step even when nb == 1 (Gap2Seq.cpp:1513)
      if (nb == 1) { pending++; cur = back[0]; }
      else { rs.skip(pending); pending = 0; cur = back[rs.next() % nb]; }
    }
    d--;
  }
  rs.skip(pending);
  if (broken) {   // "Unable to backtrace!" (Gap2Seq.cpp:1493-1510): the reference returns 0
    r.status |= G2S_ST_UNABLE_TO_BACKTRACE;
    r.count = 0;
    return;
  }
  r.left_fuz = left_fuz;
  const int from = g.left_max_fuz - left_fuz;
  r.fill_len = d0 - from;
  if (fill_out) { memcpy(fill_out, &buf[from], (size_t)r.fill_len); fill_out[r.fill_len] = 0; }
  r.count = count;
}

// Everything the host keeps of one device batch: per-gap summaries and the pack records, detached from the
// engine so that several batches can be traced later in any chosen order (the scaffold scanner needs that).
struct HostBatch {
  int k = 0;
  std::vector<g2s_gap> gaps;                  // flank pointers are not kept
  std::vector<GapOut> outs;
  std::vector<int> tier;                      // tier that produced the gap; -2 = outgrew the largest tier
  std::vector<int> pack_of;                   // index into packs
  std::vector<const char *> fill_base;        // fill arena per device pass (gaps the device traced itself, GS_TRACED)
  std::vector<const PackRec *> pack_base;     // one arena per device pass (pinned staging owned by the engine's pool,
  std::vector<std::vector<PackRec>> own;      //  or `own` when the producer has no pool)
  std::vector<std::vector<char>> own_fill;    // likewise for fills (host emulation)
  std::function<void()> on_free;              // hands the staging buffers back
  std::vector<std::pair<const void *, size_t>> staging;   // DMA targets to evict from the CPU caches when done
  std::vector<GeneralResult> gen;
  std::vector<char> buf;
  std::vector<std::shared_ptr<HostBatch>> parts;   // a batch merged from the batches of several GPUs keeps them alive
  std::vector<uint8_t> phantom;               // per gap: the caller's contains() bits disagree with the imported set (may be empty)

  HostBatch() {}
  HostBatch(const HostBatch &) = delete;
  HostBatch &operator=(const HostBatch &) = delete;
  ~HostBatch() { if (on_free) on_free(); }
  const PackRec *pack_ptr(size_t i) const { return pack_base[pack_of[i]] + outs[i].pack_off; }
  const char *fill_ptr(size_t i) const {
    return (outs[i].status & GS_TRACED) && !fill_base.empty() && fill_base[pack_of[i]] ? fill_base[pack_of[i]] + outs[i].fill_off : nullptr;
  }
  // call when the batch has been traced: the staging buffers leave the CPU caches before they are reused
  void flush_staging(ThreadPool *pool) {
    const size_t chunk = 64 << 10;
    static const long limit = getenv("G2S_FLUSH_MB") ? atol(getenv("G2S_FLUSH_MB")) : 16;   // flushing costs more than the
    size_t total = 0;                                                                       // snoops for large arenas
    for (auto &st : staging) total += st.second;
    if ((long)(total >> 20) >= limit) { staging.clear(); return; }
    for (auto &st : staging) {
      const char *base = (const char *)st.first;
      const size_t bytes = st.second, nchunk = (bytes + chunk - 1) / chunk;
      auto one = [&](size_t c) { flush_cache_range(base + c * chunk, std::min(chunk, bytes - c * chunk)); };
      if (pool) pool->parallel_for(nchunk, 1, one);
      else for (size_t c = 0; c < nchunk; c++) one(c);
    }
    staging.clear();
  }

  // SCC analysis of the general-subgraph gaps, independent per gap -> host threads
  void analyze(ThreadPool *pool) {
    const size_t n = gaps.size();
    gen.assign(n, GeneralResult());
    std::vector<uint32_t> ids;
    for (size_t i = 0; i < n; i++)
      if (tier[i] >= 0 && (outs[i].status & GS_GENERAL) && outs[i].count_fwd > 0 && !(gaps[i].flags & G2S_GAP_SKIP_CONFIDENT))
        ids.push_back((uint32_t)i);
    if (ids.empty()) return;
    // largest subgraphs first: the loop ends with the small ones
    std::sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) { return outs[a].npack > outs[b].npack; });
    auto one = [&](size_t j) { analyze_general(pack_ptr(ids[j]), outs[ids[j]].npack, gen[ids[j]]); };
    if (pool) pool->parallel_for(ids.size(), 1, one);
    else for (size_t j = 0; j < ids.size(); j++) one(j);
  }

  // what fill_gap will return / write to *right_fuz, known before the traceback (needed by the scanner to decide
  // where the next gap of the same record starts, src/Gap2Seq.cpp:392-402)
  int peek_count(size_t i) const {
    if (tier[i] == -2) return -1;
    const GapOut &o = outs[i];
    if (o.count_fwd == 0 || o.n_plen == 0) return 0;
    const bool skip = (gaps[i].flags & G2S_GAP_SKIP_CONFIDENT) != 0, all = (gaps[i].flags & G2S_GAP_ALL_PATHS) != 0;
    return (!skip && all) ? (int)o.count_all : (int)o.count_fwd;
  }
  int peek_right_fuz(size_t i) const {
    if (tier[i] == -2) return 0;
    const GapOut &o = outs[i];
    return (o.count_fwd == 0 || o.n_plen == 0) ? 0 : o.reached_fuz;
  }

  // consumes libc rand(): call in the order the reference would call fill_gap
  void trace(size_t i, g2s_result &r, char *fill_out) {
    if (fill_out) fill_out[0] = 0;
    if (tier[i] == -2) {
      memset(&r, 0, sizeof(r));
      r.count = -1; r.status = G2S_ST_MEM_LIMIT;
      return;
    }
    LibcRand rs;
    finish_one(gaps[i], outs[i], pack_ptr(i), k, tier[i], gen[i], r, fill_out, buf, rs, fill_ptr(i));
    if (i < phantom.size() && phantom[i]) r.status |= G2S_ST_PHANTOM_FLANK;
  }

  // the same from a precomputed window of the rand() stream; thread safe (scratch is the caller's)
  void trace_at(size_t i, g2s_result &r, char *fill_out, RandCursor &rc, std::vector<char> &scratch) const {
    if (fill_out) fill_out[0] = 0;
    if (tier[i] == -2) {
      memset(&r, 0, sizeof(r));
      r.count = -1; r.status = G2S_ST_MEM_LIMIT;
      return;
    }
    finish_one(gaps[i], outs[i], pack_ptr(i), k, tier[i], gen[i], r, fill_out, scratch, rc, fill_ptr(i));
    if (i < phantom.size() && phantom[i]) r.status |= G2S_ST_PHANTOM_FLANK;
  }
  // upper bound of the rand() values gap i's traceback can consume (one per step plus the length choice)
  size_t rand_bound(size_t i) const {
    if (tier[i] == -2) return 0;
    const GapOut &o = outs[i];
    if (o.count_fwd == 0 || o.n_plen == 0) return 0;
    return (size_t)std::max(o.plen[0], o.n_plen > 1 ? o.plen[1] : 0) + 2;
  }
  // Exact number of rand() values gap i's traceback consumes, if that does not depend on the random choices: one
  // for the length choice plus one per step from the path's end down to the level d_end <= left_max_fuz where the
  // walk stands on a start cell.  The device packs marked cells top level first, so the levels <= left_max_fuz are
  // the tail of the gap's records; d_end is certain when the trellis cells (PM_T) of those levels form a chain.
  bool rand_exact(size_t i, size_t *calls) const {
    *calls = 0;
    if (tier[i] == -2) return true;
    const GapOut &o = outs[i];
    if (o.count_fwd == 0 || o.n_plen == 0) return true;
    if (o.status & GS_TRACED) { *calls = o.rand_calls; return true; }
    *calls = (size_t)std::max(1, 1 + o.plen[0] - gaps[i].left_max_fuz);
    if (o.n_plen != 1 || o.start_pack[0] == NONE32) return false;
    const PackRec *pk = pack_ptr(i);
    const int lmf = gaps[i].left_max_fuz;
    if (o.plen[0] <= lmf) return false;
    uint32_t n_t = 0, the_one = NONE32;   // trellis cells on level lmf
    for (uint32_t q = o.npack; q-- > 0;) {   // tail first: levels ascend
      const PackRec &r = pk[q];
      if ((int)r.level > lmf) break;
      if (!(r.meta & PM_T)) continue;
      if ((int)r.level == lmf) { n_t++; the_one = q; }
    }
    // the walk stops on level lmf iff its cell there is a start; with one trellis cell on that level that is certain
    return n_t == 1 && (pk[the_one].meta & PM_START) != 0;
  }
  size_t fill_bytes(size_t i) const {
    const g2s_gap &g = gaps[i];
    return (size_t)g.gap_len + k + g.gap_err + g.left_max_fuz + g.right_max_fuz + 3;
  }
};

// Tracebacks of many gaps, in the order the reference would call fill_gap, consuming libc rand() exactly like
// successive calls would - but on all host threads.  Gap j's traceback starts at the stream position where gap
// j-1's stopped, and how many values a traceback consumes is usually known without running it: one per level between
// the path's end and the level where it meets the left start, which HostBatch::rand_exact() reads off the bottom of
// the trellis (certain when the trellis is a single chain down there and there is one path length).  So the gaps are
// traced in rounds: a round covers the gaps up to and including the next one whose consumption is not certain, each
// at its position in a window of the stream computed ahead (RandAhead), on all threads; the round's last gap tells
// the true position for the next round.  Consumption is verified after every round, so a wrong "certain" could only
// cost a re-trace, never a wrong result.  The outcome equals tracing gap after gap from libc.
struct TraceJob { HostBatch *hb; size_t idx; };
struct TraceStats { std::atomic<unsigned long long> calls{0}, parallel_calls{0}, rounds{0}, retraced{0}; };
inline TraceStats &trace_stats() { static TraceStats s; return s; }   // observability for the tests

// `direct` (optional): fills the device traced itself are not copied into the arena; direct[j] points at the gap's fill
// wherever it lives (the pinned staging of its batch, or the arena) - for callers that only read the fills once.
inline void trace_many(const std::vector<TraceJob> &jobs, std::vector<g2s_result> &res, RawBuf<char> &arena_buf,
                       std::vector<size_t> &fill_off, ThreadPool *pool, RandAhead *ahead,
                       std::vector<const char *> *direct = nullptr) {
  const size_t n = jobs.size();
  res.resize(n);
  fill_off.resize(n + 1);
  size_t tot = 0;
  for (size_t j = 0; j < n; j++) { fill_off[j] = tot; tot += jobs[j].hb->fill_bytes(jobs[j].idx); }
  fill_off[n] = tot;
  char *arena = arena_buf.ensure(tot + 1);
  if (direct) {
    direct->resize(n);
    for (size_t j = 0; j < n; j++) { const char *d = jobs[j].hb->fill_ptr(jobs[j].idx); (*direct)[j] = d ? d : &arena[fill_off[j]]; }
  }
  auto dst_of = [&](size_t j) -> char * { return direct && (*direct)[j] != &arena[fill_off[j]] ? nullptr : &arena[fill_off[j]]; };
  RandAhead local;
  if (!ahead) { ahead = &local; if (pool && n >= 4) local.begin(); }
  trace_stats().calls++;
  if (!pool || n < 4 || !ahead->ok()) {   // sequential, straight from libc
    if (ahead->ok()) ahead->commit(0);
    for (size_t j = 0; j < n; j++) jobs[j].hb->trace(jobs[j].idx, res[j], dst_of(j));
    return;
  }
  std::vector<size_t> pred(n), act(n), off(n + 1);
  std::vector<uint8_t> certain(n);
  std::vector<size_t> bounds(n);
  pool->parallel_for(n, 16, [&](size_t j) {
    const HostBatch &hb = *jobs[j].hb;
    bounds[j] = hb.rand_bound(jobs[j].idx);
    certain[j] = hb.rand_exact(jobs[j].idx, &pred[j]) ? 1 : 0;
  });
  size_t bound = 0;
  for (size_t j = 0; j < n; j++) bound += bounds[j];
  ahead->extend(bound + 1);
  const uint32_t *vals = ahead->data();
  size_t final_upto = 0, base = 0;
  trace_stats().parallel_calls++;
  auto trace_one = [&](size_t j, std::vector<char> &scratch) {
    RandCursor rc{vals, off[j], ahead};
    jobs[j].hb->trace_at(jobs[j].idx, res[j], dst_of(j), rc, scratch);
    act[j] = rc.pos - off[j];
  };
  // Segments: the gaps up to and including the next uncertain one.  A segment's stream position is known once the
  // previous segment's uncertain gap has been traced, so those gaps form the only sequential chain; they are handed
  // out first (work items are taken in this order) and the certain gaps fill the other threads in the meantime.
  {
    std::vector<size_t> seg_of(n), seg_first, seg_last, order;
    order.reserve(n);
    size_t sidx = 0;
    for (size_t lo = 0; lo < n;) {
      size_t hi = lo;
      while (hi + 1 < n && certain[hi]) hi++;
      seg_first.push_back(lo);
      seg_last.push_back(hi);
      order.push_back(hi);
      for (size_t j = lo; j < hi; j++) order.push_back(j);
      for (size_t j = lo; j <= hi; j++) seg_of[j] = sidx;
      sidx++;
      lo = hi + 1;
    }
    const size_t nseg = seg_first.size();
    std::vector<size_t> seg_base(nseg + 1, 0);
    std::atomic<size_t> ready{1};          // segments whose stream position is known
    trace_stats().rounds += nseg;
    pool->parallel_for(n, 1, [&](size_t q) {
      static thread_local std::vector<char> scratch;
      const size_t j = order[q], sg = seg_of[j];
      for (unsigned spins = 0; ready.load(std::memory_order_acquire) <= sg; spins++)   // the chain of uncertain gaps is
        if (spins > 2000) std::this_thread::yield();                                   // being traced by another thread
      size_t o = seg_base[sg];
      for (size_t i = seg_first[sg]; i < j; i++) o += pred[i];
      off[j] = o;
      trace_one(j, scratch);
      if (j == seg_last[sg]) {   // the segment's last gap: it fixes where the next segment starts
        seg_base[sg + 1] = o + act[j];
        ready.store(sg + 2, std::memory_order_release);
      }
    });
    // every certain gap must have consumed what was predicted; otherwise everything behind it is traced again
    size_t bad = n;
    for (size_t j = 0; j < n && bad == n; j++) if (act[j] != pred[j] && certain[j]) bad = j;
    if (bad == n) { final_upto = n; base = seg_base[nseg]; }
    else { final_upto = bad + 1; base = off[bad] + act[bad]; trace_stats().retraced += n - bad - 1; }
  }
  std::vector<char> my_scratch;
  while (final_upto < n) {   // only after a wrong prediction: rounds up to the next uncertain gap
    const size_t lo = final_upto;
    size_t hi = lo;
    while (hi + 1 < n && certain[hi]) hi++;
    off[lo] = base;
    for (size_t j = lo; j < hi; j++) off[j + 1] = off[j] + pred[j];
    for (size_t j = lo; j <= hi; j++) trace_one(j, my_scratch);
    size_t bad = hi;
    for (size_t j = lo; j < hi; j++) if (act[j] != pred[j]) { bad = j; break; }
    base = off[bad] + act[bad];
    final_upto = bad + 1;
  }
  ahead->commit(base);
}


}  // namespace g2s
