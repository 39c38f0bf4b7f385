// g2s_chain.cuh - the chain tier: gaps of Gap2Seq::fill_gap (reference src/Gap2Seq.cpp:858-1556) whose borders never
// hold more than ONE state - the right pass walks back along a path of in-degree-1 nodes, the forward pass walks along a
// path of out-degree-1 nodes - which is what a gap in unique sequence covered by clean reads looks like.
//
// For such a gap every table of the reference degenerates: the right set is a string, the forward table is a string,
// counts are 1, the path subgraph is a path, the traceback has no choice.  The tier therefore keeps NO per-gap hash table
// and no cell arrays.  It runs in two kernels:
//
//   k_chain_walk    one LANE per walk (two per gap: forward from the left flank up to the first pruned level, backward
//                   from the right flank; the forward chain's levels beyond that ARE the backward chain, see the finish).  A walk
//                   is the chain of dependent gathers in the global k-mer table and nothing else: every gather returns the
//                   skip hint of the tip (the next <= 40 bases of its unique extension), the lane appends those bases to
//                   the walk's 2-bit packed string in HBM and jumps to the node the hint ends at.  Thousands of walks are
//                   in flight per SM, so the pass is bound by the longest gap's gather chain, not by per-level work.
//   k_chain_finish  one WARP per gap.  Reads the two strings and derives, level-parallel, everything fill_gap computes:
//                   where the forward chain meets the right set (the meet-in-the-middle prune of Gap2Seq.cpp:1050), the
//                   first-found check (Gap2Seq.cpp:1107-1159), the backward subgraph and its statistics
//                   (Gap2Seq.cpp:1173-1435), the safe/unsafe case mask and the fill string (Gap2Seq.cpp:1437-1517; with one
//                   trellis cell per level the random traceback has no choice to make, only rand() calls to count).
//
// The tier proves its own applicability.  Whenever anything could make the reference's result differ from the chain
// reading - a border that would hold two states, a start k-mer that is not on the chain, a k-mer met twice (either
// strand), two accepted path lengths, a forward state beyond the right flank that is in the right set again - the gap
// is handed on (GS_NOT_CHAIN) and the thin tier runs it with the full tables.  Results of the gaps it keeps are
// bit-identical to the thin tier's (tests/test_chain_emul.py on the CPU, the -m gpu parity tests on the device).
#pragma once
#include "g2s_fill.cuh"

namespace g2s {

enum { GS_NOT_CHAIN = 64 };   // GapOut.status: the chain tier declined the gap

static const uint32_t CHAIN_END_LIMIT = 0xFFu;      // ChainWalk.end_mask: the walk reached its level limit
static const uint32_t CHAIN_END_INTERNAL = 0x100u;  // the table contradicts its own adjacency bytes

struct ChainWalk {      // what a walk left behind, 16 B
  uint32_t n;           // states recorded (the string holds k + n - 1 bases); 0: the start k-mer is not in the graph
  uint32_t end_mask;    // bits 0-15: out-mask of the last state in walk orientation (0: dead end, several bits: branch), or
                        // CHAIN_END_*; bits 24-27 (backward walk): out-mask of the START state against the walk, i.e. the
                        // successors of the last right start in forward orientation
  uint32_t lookups;     // gathers in the global table
  uint32_t ok_upto;     // backward walk: states 1 .. ok_upto each have exactly ONE neighbour against the walk direction (the
                        // skip hint of the other orientation, read at every gather, covers the segment walked before it)
};
G2S_HD uint32_t chain_end(const ChainWalk &w) { return w.end_mask & 0xFFFFu; }

struct ChainParams {
  Table table;
  const Table *tables;        // per-gap graphs (FillParams)
  const uint32_t *graph_of;
  const char *flanks;
  const GapDev *gaps;
  const uint32_t *order;      // gap ids of this pass
  uint32_t n_gaps;
  GapOut *outs;
  uint64_t *strs;             // per position in `order`: f_words words of the forward string, then r_words of the right one
  uint32_t f_words, r_words;
  uint32_t walk_stride;       // every walk_stride-th lane of the walk kernel carries a walk (32 / walk_stride walks per warp)
  ChainWalk *walks;           // [2q] forward, [2q + 1] right
  char *fills;                // fill arena
  unsigned long long fill_cap;
  unsigned long long *fill_cursor;
};

// words a walk of up to `levels` levels needs: k + levels bases, one partial word, one word of slack for str_kmer
inline uint32_t chain_words(int k, int levels) { return (uint32_t)((2 * ((size_t)k + levels) + 63) / 64 + 2); }

// ------------------------------------------------------------------------------------------------ packed strings
// Base p of a string lives in word p / 32, bits 63 - 2 (p % 32) .. 62 - 2 (p % 32): first base highest, like a k-mer.
struct BaseWriter {
  uint64_t *out;
  uint64_t acc;
  uint32_t nb, w;       // bits held in acc (from the top), words written
  G2S_D void init(uint64_t *p) { out = p; acc = 0; nb = 0; w = 0; }
  G2S_D void put(uint64_t v, uint32_t n) {   // the n (1..64) low bits of v
    if (n < 64) v &= (1ULL << n) - 1;
    const uint32_t room = 64 - nb;
    if (n < room) { acc |= v << (room - n); nb += n; return; }
    const uint32_t rem = n - room;             // 0..63
    out[w++] = acc | (v >> rem);
    acc = rem ? v << (64 - rem) : 0;
    nb = rem;
  }
  G2S_D void flush() { out[w] = acc; out[w + 1] = 0; }
};

// The finish kernel only reads the strings (the walk kernel wrote them): plain cached loads, a warp's 32 consecutive
// levels touch two or three words.
G2S_D uint64_t ld_ro(const uint64_t *p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
G2S_D uint64_t str_kmer(const uint64_t *W, uint32_t p, int k) {   // bases p .. p + k - 1 as a k-mer
  const uint32_t q = p >> 5, s = 2 * (p & 31);
  const uint64_t hi = ld_ro(&W[q]), lo = ld_ro(&W[q + 1]);
  const uint64_t v = s ? (hi << s) | (lo >> (64 - s)) : hi;
  return v >> (64 - 2 * k);
}
G2S_D uint32_t str_base(const uint64_t *W, uint32_t p) { return (uint32_t)(ld_ro(&W[p >> 5]) >> (62 - 2 * (p & 31))) & 3u; }

// ------------------------------------------------------------------------------------------------ the walk
// The walk's lookup: the home slot and its neighbour are requested together, so that a key one probe away from home
// (a third of them at the table's load) does not cost a second dependent round trip.
struct ChainProbe {     // the two requests of a lookup in flight
  uint64_t h, h1, ak, av, a0, a1, bk, bv, b0, b1;
};
G2S_D void chain_probe_issue(const Table &t, uint64_t canon, ChainProbe &p) {
  p.h = mix64(canon) & t.cap_mask;
  p.h1 = (p.h + 1) & t.cap_mask;
#ifdef __CUDA_ARCH__
  // one statement: both requests leave before either result is looked at
  asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%8];\n\tld.global.nc.v4.u64 {%4, %5, %6, %7}, [%9];"
               : "=l"(p.ak), "=l"(p.av), "=l"(p.a0), "=l"(p.a1), "=l"(p.bk), "=l"(p.bv), "=l"(p.b0), "=l"(p.b1)
               : "l"(t.slots + p.h), "l"(t.slots + p.h1));
#else
  const Slot a = t.slots[p.h], b = t.slots[p.h1];
  p.ak = a.key; p.av = a.val; p.a0 = a.h[0]; p.a1 = a.h[1];
  p.bk = b.key; p.bv = b.val; p.b0 = b.h[0]; p.b1 = b.h[1];
#endif
}
G2S_D bool chain_probe_resolve(const Table &t, uint64_t canon, const ChainProbe &p, Slot *out) {
#ifdef __CUDA_ARCH__
  // straight-line selects: the compiler must not sink the second request behind a test of the first result
  uint64_t sv, s0, s1;
  unsigned hit;
  asm volatile("{\n\t.reg .pred pa, pb, pe;\n\t"
               "setp.eq.u64 pa, %4, %10;\n\tsetp.eq.u64 pb, %7, %10;\n\tsetp.ne.u64 pe, %4, %11;\n\t"
               "and.pred pb, pb, pe;\n\t"
               "selp.b64 %0, %5, %8, pa;\n\tselp.b64 %1, %6, %9, pa;\n\t"
               "selp.b64 %2, %12, %13, pa;\n\t"
               "or.pred pa, pa, pb;\n\tselp.u32 %3, 1, 0, pa;\n\t}"
               : "=l"(sv), "=l"(s0), "=l"(s1), "=r"(hit)
               : "l"(p.ak), "l"(p.av), "l"(p.a0), "l"(p.bk), "l"(p.bv), "l"(p.b0), "l"(canon), "l"(KEY_EMPTY), "l"(p.a1), "l"(p.b1));
  if (hit) { out->key = canon; out->val = sv; out->h[0] = s0; out->h[1] = s1; return true; }
#else
  if (p.ak == canon) { out->key = canon; out->val = p.av; out->h[0] = p.a0; out->h[1] = p.a1; return true; }
  if (p.ak != KEY_EMPTY && p.bk == canon) { out->key = canon; out->val = p.bv; out->h[0] = p.b0; out->h[1] = p.b1; return true; }
#endif
  if (p.ak == KEY_EMPTY || p.bk == KEY_EMPTY) return false;
  uint64_t h = (p.h1 + 1) & t.cap_mask;
  for (;;) {
    const Slot s = load_slot32(t.slots + h);
    if (s.key == canon) { *out = s; return true; }
    if (s.key == KEY_EMPTY) return false;
    h = (h + 1) & t.cap_mask;
  }
}
G2S_D bool chain_find_slot(const Table &t, uint64_t canon, Slot *out) {
  ChainProbe p;
  chain_probe_issue(t, canon, p);
  return chain_probe_resolve(t, canon, p, out);
}

// One chain of dependent gathers: from x0 (walk orientation) along unique successors for at most `maxlev` levels.
// backcheck (the backward walk): every gather also reads the tip's hint of the OTHER orientation - the length of the
// unique extension against the walk - and thereby proves that the nodes of the segment just walked have one neighbour
// in that direction too (that is what lets the forward chain be read off the backward one); the final state is gathered
// for that purpose even when the level limit ends the walk.
G2S_D void chain_walk(const Table &T, uint64_t x0, int maxlev, uint64_t *W, ChainWalk *out, bool backcheck) {
  const int k = T.k;
  ChainWalk r{0, 0, 1, 0};
  Canon c = canonical(x0, k);
  Slot sl;
  if (!chain_find_slot(T, c.key, &sl)) { *out = r; return; }
  BaseWriter bw;
  bw.init(W);
  bw.put(x0, (uint32_t)(2 * k));
  uint64_t x = x0;
  int lev = 0;
  uint32_t endm = CHAIN_END_LIMIT;
  const uint32_t aux = backcheck ? out_mask((uint32_t)(sl.val & 0xFF), c.strand ^ 1u) : 0u;
  bool ok = true;
  while (lev < maxlev) {
    const uint32_t mask = out_mask((uint32_t)(sl.val & 0xFF), c.strand);
    if (mask == 0 || (mask & (mask - 1))) { endm = mask; break; }
    Hint h = hint_of(sl, c.strand);
    if (h.len == 0) { h.lo = mask == 1 ? 0 : mask == 2 ? 1 : mask == 4 ? 2 : 3; h.hi = 0; h.len = 1; }   // a table without hints
    const uint32_t j = h.len < (uint32_t)(maxlev - lev) ? h.len : (uint32_t)(maxlev - lev);
    const uint32_t sh = 2 * (h.len - j);                 // the first j bases = bases >> sh
    // the next lookup leaves first: writing this segment's bases overlaps its latency
    x = hint_state(x, h, j, T.kmask);
    lev += (int)j;
    const bool last = lev >= maxlev;
    const bool look = !last || backcheck;
    ChainProbe pr;
    if (look) { c = canonical(x, k); chain_probe_issue(T, c.key, pr); }
    if (2 * j > 64) bw.put(sh < 16 ? (uint64_t)(h.hi >> sh) : 0, 2 * j - 64);
    bw.put(hint_shr(h, sh), 2 * j > 64 ? 64u : 2 * j);
    if (!look) break;
    r.lookups++;
    if (!chain_probe_resolve(T, c.key, pr, &sl)) { endm = CHAIN_END_INTERNAL; break; }
    if (backcheck) {
      if (ok && hint_of(sl, c.strand ^ 1u).len >= j) r.ok_upto = (uint32_t)lev;
      else ok = false;
    }
    if (last) break;
  }
  bw.flush();
  r.n = (uint32_t)lev + 1;
  r.end_mask = endm | (aux << 24);
  *out = r;
}

// ------------------------------------------------------------------------------------------------ the finish
// G: the thread group of one gap (a warp on the device, one thread in the host emulation) with
//    rank(), size(), any(bool), min_i32(int), max_i32(int), add_u32(uint32_t), bcast_u64(v) (rank 0's value).
// tk: group-shared scratch of 4 * (MAX_FUZ + 1) words: the flank k-mers, parsed once - [j] right-flank target j,
//     [32 + d] left start of level d, [64 + o] right start at offset o of the start string, [96 + j] reverse complement of
//     target j (what a target looks like in the right chain's string).

// is x a node of the right chain (either strand)?
template <class G>
G2S_D bool chain_member(const uint64_t *WR, int nR, uint64_t x, int k) {
  const uint64_t rx = revcomp(x, k);
  bool hit = false;
  for (int i = G::rank(); i < nR; i += G::size()) {
    const uint64_t y = str_kmer(WR, (uint32_t)i, k);
    if (y == x || y == rx) hit = true;
  }
  return G::any(hit);
}

template <class G>
G2S_D void chain_finish(const ChainParams &P, const GapDev &gd, const uint64_t *WF, const uint64_t *WR, const ChainWalk wf,
                        const ChainWalk wr, uint64_t *tk, GapOut *out) {
  const Table T = gap_table(P.table, P.tables, P.graph_of, gd.id);
  const int k = T.k, rank = G::rank(), gsz = G::size();
  const int g = gd.gap_len, e = gd.gap_err, lmf = gd.lmf, rmf = gd.rmf;
  const bool all_paths = (gd.flags & 2) != 0, skip_confident = (gd.flags & 1) != 0;
  // the walks were cut at right_half = rmf + ceil((g + e) / 2) and Dmax = g + e + lmf + rmf levels (Gap2Seq.cpp:862-863)
  const int pf = g / 2 + e / 2 + lmf;               // first pruned level, Gap2Seq.cpp:1050
  const char *lstr = P.flanks + gd.flank_off, *rstr = lstr + k + lmf, *rsstr = P.flanks + gd.rstart_off;
  const int nF = (int)wf.n, nR = (int)wr.n, cw = nF - 1;
  const int NONE = 0x7fffffff;

  uint64_t *lkk = tk + (MAX_FUZ + 1), *rsk = tk + 2 * (MAX_FUZ + 1), *rtk = tk + 3 * (MAX_FUZ + 1);
  for (int j = rank; j <= MAX_FUZ; j += gsz) {
    if (j <= rmf) { tk[j] = kmer_from_ascii(rstr + j, k); rsk[j] = kmer_from_ascii(rsstr + j, k); rtk[j] = revcomp(tk[j], k); }
    if (j <= lmf) lkk[j] = kmer_from_ascii(lstr + j, k);
  }
  G::sync();
  bool internal = ((chain_end(wf) | chain_end(wr)) & CHAIN_END_INTERNAL) != 0;
  bool fb = nF == 0 || nR == 0 || internal;         // a start k-mer that is absent: the staggered starts decide, not this tier
  const bool r_branch = chain_end(wr) != CHAIN_END_LIMIT && (chain_end(wr) & (chain_end(wr) - 1)) != 0;
  const bool f_branch = chain_end(wf) != CHAIN_END_LIMIT && (chain_end(wf) & (chain_end(wf) - 1)) != 0;
  // the right pass never prunes: a state with two predecessors below the level limit makes a border of two
  if (r_branch) fb = true;
  // the forward chain must cover the levels that inject starts
  if (cw < lmf) fb = true;

  // ---- staggered starts (Gap2Seq.cpp:953-968, 1082-1105): each one is either the chain's own state of its level (the
  // injection changes nothing: same node, same strand, count already 1) or absent from the graph
  if (!fb) {
    bool bad = false;
    for (int c = 1 + rank; c <= rmf; c += gsz) {            // right start of iteration c = state of index c (rmf <= right_half)
      const uint64_t s = ld_sh(&rsk[rmf - c]);
      if (c < nR && revcomp(s, k) == str_kmer(WR, (uint32_t)c, k)) continue;
      uint32_t adj;
      if (table_find(T, canonical(s, k).key, &adj) >= 0) bad = true;
    }
    for (int d = 1 + rank; d <= lmf; d += gsz) {
      const uint64_t s = ld_sh(&lkk[d]);
      if (s == str_kmer(WF, (uint32_t)d, k)) continue;
      uint32_t adj;
      if (table_find(T, canonical(s, k).key, &adj) >= 0) bad = true;
    }
    if (G::any(bad)) fb = true;
  }

  // ---- where the forward chain ends.  Below pf nothing is pruned; the forward walk covers levels 0 .. pf.  From pf on a
  // state survives iff its node is in the right set.  The right chain R[0], R[1], .. walks UNIQUE predecessors from R[0] =
  // the last right start, and its gathers proved that R[1] .. R[ok_upto] have one SUCCESSOR each as well.  So if the
  // forward chain stands on R[i0] at level pf, it walks R[i0 - 1], .. , R[0] on levels pf + 1 .. L = pf + i0 and nothing
  // else can enter the border there: those levels need neither a walk nor a lookup - the forward chain beyond pf is the
  // right chain read backwards.  What comes behind R[0] (one successor, from the start slot's adjacency) is settled by one
  // membership scan.
  int top = -1, L = NONE;
  if (!fb) {
    if (cw < pf) top = cw;                    // the chain ended (dead end or branch) before the first pruned level
    else {
      const uint64_t fp = str_kmer(WF, (uint32_t)pf, k), want = revcomp(fp, k);   // the right string holds reverse complements
      // what follows R[0] on the forward chain, if anything: its one successor (several: each would need its own test)
      const uint32_t fwd0 = (wr.end_mask >> 24) & 15u;
      const bool has_next = fwd0 != 0 && !(fwd0 & (fwd0 - 1));
      const uint64_t nxt = has_next ? succ_kmer(revcomp(str_kmer(WR, 0, k), k), fwd0 == 1 ? 0 : fwd0 == 2 ? 1 : fwd0 == 4 ? 2 : 3, T) : 0;
      const uint64_t rnxt = revcomp(nxt, k);
      int i0 = NONE;
      bool flipped = false, next_inside = false;
      for (int i = rank; i < nR; i += gsz) {     // ONE pass over the right chain answers all three questions
        const uint64_t y = str_kmer(WR, (uint32_t)i, k);
        if (y == want) { i0 = i < i0 ? i : i0; }
        if (y == fp) flipped = true;             // the node is in the right set on its other strand: not a chain to read off
        if (has_next && (y == nxt || y == rnxt)) next_inside = true;
      }
      i0 = G::min_i32(i0);
      if (G::any(flipped)) fb = true;
      next_inside = G::any(next_inside);
      if (i0 == NONE) top = pf - 1;              // not in the right set: the border dies on level pf
      else {
        if ((uint32_t)i0 > wr.ok_upto) fb = true;        // a node on the way to R[0] has a second successor
        L = pf + i0;
        top = L;
        if (L < g + e + lmf + rmf) {                     // a level behind L exists (L < Dmax)
          if (fwd0 & (fwd0 - 1)) fb = true;
          else if (next_inside) fb = true;               // the chain would go on inside the right set
        }
      }
    }
    if (f_branch && top == cw) fb = true;     // the branching state is expanded: a border of two
    if (top < lmf) fb = true;                 // a start level without a chain state: injection starts a new chain
  }
  // state and last base of forward level c <= top: from the forward string up to pf, from the right string beyond
  auto f_at = [&](int c) -> uint64_t {
    return c <= cw ? str_kmer(WF, (uint32_t)c, k) : revcomp(str_kmer(WR, (uint32_t)(L - c), k), k);
  };
  auto f_base = [&](int c) -> uint32_t {
    return c <= cw ? str_base(WF, (uint32_t)(c + k - 1)) : (str_base(WR, (uint32_t)(L - c)) ^ 2u);
  };

  // ---- one pass over the chain's states: no node twice (either strand), target hits, all-paths seeds.
  // A chain of forced steps that meets a node again is periodic from there on (same strand) or retraces its own reverse
  // complement (other strand); either way its LAST state, or the reverse complement of its FIRST, shows up in the chain -
  // so three comparisons per level prove that all nodes are distinct.
  uint32_t best = 0xFFFFFFFFu;   // (err << 6 | j) of the first-found check
  uint32_t n_sink = 0;
  int min_sink = NONE, sB = -1;
  if (!fb) {
    const uint64_t F0 = str_kmer(WF, 0, k), Ftop = f_at(top), rcFtop = revcomp(Ftop, k), rcF0 = revcomp(F0, k);
    const int jt = rmf - 1, dlo = lmf + g - e > 0 ? lmf + g - e : 0;
    bool rep = false;
    // a level can hit a target only within e of the target's nominal level (and seeds start at dlo >= that bound); a
    // 64-bit signature of the targets spares most of those levels the rmf + 1 comparisons
    const int hit_lo = g + lmf - e > 0 ? g + lmf - e : 0;
    // (two 64-bit signatures from different bits of one product: ~3 % of the levels pass both by chance)
    uint64_t sig = 0, rsig = 0, sig2 = 0, rsig2 = 0;
    for (int j = 0; j <= rmf; j++) {
      const uint64_t ha = ld_sh(&tk[j]) * 0x9E3779B97F4A7C15ULL, hb = ld_sh(&rtk[j]) * 0x9E3779B97F4A7C15ULL;
      sig |= 1ULL << (ha >> 58); sig2 |= 1ULL << ((ha >> 52) & 63);
      rsig |= 1ULL << (hb >> 58); rsig2 |= 1ULL << ((hb >> 52) & 63);
    }
    // levels 0 .. min(top, cw) come from the forward string; levels cw + 1 .. top are the states R[L - c], read from the
    // right string, whose k-mers are reverse complements - there the constants are turned over instead of every state.
    // One loop over both ranges: lane steps of 32 levels, a level is 2 loads, a funnel shift and a handful of compares.
    const int hiF = top < cw ? top : cw;
    const uint64_t c_top = Ftop, c_rtop = rcFtop, c_rf0 = rcF0;
    for (int c = rank; c <= top; c += gsz) {
      const bool fwd = c <= hiF;
      const uint64_t x = fwd ? str_kmer(WF, (uint32_t)c, k) : str_kmer(WR, (uint32_t)(L - c), k);   // fwd: F[c], else rc(F[c])
      if (fwd ? ((c < top && x == c_top) || x == c_rtop || x == c_rf0) : ((c < top && x == c_rtop) || x == c_top || x == F0)) rep = true;
      const uint64_t hx = x * 0x9E3779B97F4A7C15ULL;
      if (c < hit_lo || !(((fwd ? sig : rsig) >> (hx >> 58)) & ((fwd ? sig2 : rsig2) >> ((hx >> 52) & 63)) & 1)) continue;
      const uint64_t *tt = fwd ? tk : rtk;
      for (int j = 0; j <= rmf; j++) {
        if (x != ld_sh(&tt[j])) continue;
        const int err = c > g + lmf + j ? c - (g + lmf + j) : g + lmf + j - c;
        if (err <= e) { const uint32_t key = ((uint32_t)err << 6) | (uint32_t)j; best = key < best ? key : best; }
        if (j == jt && c >= dlo) { n_sink++; min_sink = c < min_sink ? c : min_sink; sB = c > sB ? c : sB; }
      }
    }
    if (G::any(rep)) fb = true;
  }
  // group-wide: the smallest (err, j) key (err < 2^25), the seeds
  {
    int b = best == 0xFFFFFFFFu ? NONE : (int)best;
    b = G::min_i32(b);
    best = b == NONE ? 0xFFFFFFFFu : (uint32_t)b;
    n_sink = G::add_u32(n_sink);
    min_sink = G::min_i32(min_sink);
    sB = G::max_i32(sB);
  }

  bool found = false;
  int d0 = -1, J = 0, c_found = 0;
  if (!fb && best != 0xFFFFFFFFu) {
    const int err = (int)(best >> 6);
    J = (int)(best & 63u);
    const int len1 = g + lmf + J + err, len2 = g + lmf + J - err;
    const uint64_t t = ld_sh(&tk[J]);
    const bool h1 = len1 <= top && f_at(len1) == t;
    const bool h2 = err > 0 && len2 >= 0 && len2 <= top && f_at(len2) == t;
    if (h1 && h2) fb = true;                  // two accepted lengths: rand() picks one (Gap2Seq.cpp:1440)
    found = true;
    d0 = h1 ? len1 : len2;
    c_found = g + lmf + rmf + err;
  }

  // ---- backward subgraph (Gap2Seq.cpp:1173-1312) of a chain: its levels from the highest seed down to the first start
  // cell met on the way, one vertex per level; the trellis of the traceback likewise from the accepted target
  uint32_t nB = 0, n_edges = 0, n_src = 0;
  int max_src = -1, dS = -1, dT = -1;
  if (!fb && found) {
    if (skip_confident || (all_paths && rmf < 1)) { n_sink = 0; min_sink = NONE; sB = -1; }
    else if (!all_paths) { n_sink = 1; min_sink = d0; sB = d0; }                  // -best-only: the accepted target is the seed
    // start cells: d <= lmf whose state is the left start of that level (checked above: the only way to share its node)
    int s1 = -1, s2 = -1;
    for (int d = rank; d <= lmf; d += gsz) {
      if (ld_sh(&lkk[d]) != str_kmer(WF, (uint32_t)d, k)) continue;
      if (d <= sB) s1 = d;
      if (d <= d0) s2 = d;
    }
    dS = G::max_i32(s1);
    dT = G::max_i32(s2);
    if (dT < 0) fb = true;                    // cannot happen (level 0 is a start): "Unable to backtrace" is the host's to say
    if (n_sink > 0) {
      if (dS < 0) fb = true;
      nB = (uint32_t)(sB - dS + 1); n_edges = (uint32_t)(sB - dS); n_src = 1; max_src = dS;
    }
  }

  // ---- the fill (Gap2Seq.cpp:1437-1517 with one trellis cell per level): bases of levels dT+1 .. d0, upper case within k
  // levels below a safe vertex or the path's end
  unsigned long long fill_off = 0;
  if (!fb && found) {
    const unsigned long long need = ((unsigned long long)(d0 - dT) + 1 + 7) & ~7ULL;   // multiples of 8: aligned 8-byte stores
    unsigned long long off = 0;
    if (rank == 0) off = a_add64(P.fill_cursor, need);
    off = G::bcast_u64(off);
    if (P.fills == nullptr || off + need > P.fill_cap) fb = true;
    else {
      fill_off = off;
      char *outc = P.fills + off;
      // a vertex of the path subgraph (levels dS .. sB, alone on its level) is safe iff no sink edge jumps over it, i.e.
      // up to the lowest seed; a trellis cell that is no vertex gets the sink's verdict (node2boost default):
      // safe whenever seeds exist.  So the unsafe levels are exactly uLo .. sB, and a base is lower case iff the k levels
      // from its own upwards all lie in there.
      const int uLo = n_sink == 0 ? 0 : (dS > min_sink + 1 ? dS : min_sink + 1);
      for (int d = dT + 1 + rank; d <= d0; d += gsz) {
        const int wend = d + k - 1 < d0 ? d + k - 1 : d0;
        const bool up = d0 - d < k || skip_confident || (n_sink > 0 && !(d >= uLo && wend <= sB));
        // 'A' 'C' 'T' 'G' (GATB's 2-bit order) as the bytes of one word, lower case = bit 5: no branches
        outc[d - 1 - dT] = (char)(((0x47544341u >> (8 * f_base(d))) & 0xFFu) | (up ? 0u : 0x20u));
      }
      if (rank == 0) outc[d0 - dT] = 0;
    }
  }

  if (rank == 0) {
    GapOut o;
    memset(&o, 0, sizeof(o));
    o.plen[0] = o.plen[1] = -1;
    o.start_pack[0] = o.start_pack[1] = NONE32;
    o.max_src_level = -1; o.min_sink_level = NONE;
    o.table_lookups = wf.lookups + wr.lookups;
    if (fb) {
      o.status = GS_NOT_CHAIN | GS_DONE | (internal ? GS_INTERNAL : 0);
    } else {
      o.status = GS_DONE;
      o.cells_right = (uint32_t)nR;
      const int last = (!all_paths && found && c_found < top) ? c_found : top;   // -best-only ends the table at the level of the hit
      o.cells_left = (uint32_t)(last + 1);
      o.top_level = (uint32_t)last;
      o.levels_r = (uint32_t)nR; o.levels_f = (uint32_t)last;
      int shared = 0;                                                           // levels that stand on a node of the right chain
      if (L != NONE) { const int hi = L < last ? L : last, lo = L - nR + 1 > 0 ? L - nR + 1 : 0; if (hi >= lo) shared = hi - lo + 1; }
      o.nodes = (uint32_t)(nR + last + 1 - shared);
      if (found) {
        o.status |= GS_TRACED;
        o.count_fwd = 1; o.n_plen = 1; o.plen[0] = d0; o.reached_fuz = J;
        o.count_all = n_sink > MAX_PATHS ? MAX_PATHS : n_sink;
        o.nB = nB; o.n_edges = n_edges; o.n_src = n_src; o.n_sink = n_sink;
        o.max_src_level = max_src; o.min_sink_level = min_sink;
        o.fill_off = fill_off; o.fill_len = (uint32_t)(d0 - dT); o.left_fuz = lmf - dT;
        o.rand_calls = 1u + o.fill_len;
      }
    }
    *out = o;
  }
}

}  // namespace g2s
