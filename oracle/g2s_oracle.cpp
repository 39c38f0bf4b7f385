// See g2s_oracle.hpp.  TEST INFRASTRUCTURE ONLY (parity checker / CPU baseline).
#include "g2s_oracle.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <unordered_set>

namespace g2s_oracle {

/* ------------------------------------------------------------------ graph */

static const uint64_t EMPTY = ~0ULL;

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

// GATB tools/math/LargeInt1.pri:136-154
static inline uint64_t revcomp64(uint64_t x, int k) {
  uint64_t r = x;
  r = ((r >> 2 & 0x3333333333333333ULL) | (r & 0x3333333333333333ULL) << 2);
  r = ((r >> 4 & 0x0F0F0F0F0F0F0F0FULL) | (r & 0x0F0F0F0F0F0F0F0FULL) << 4);
  r = ((r >> 8 & 0x00FF00FF00FF00FFULL) | (r & 0x00FF00FF00FF00FFULL) << 8);
  r = ((r >> 16 & 0x0000FFFF0000FFFFULL) | (r & 0x0000FFFF0000FFFFULL) << 16);
  r = ((r >> 32 & 0x00000000FFFFFFFFULL) | (r & 0x00000000FFFFFFFFULL) << 32);
  r ^= 0xAAAAAAAAAAAAAAAAULL;
  return r >> (2 * (32 - k));
}

// GATB tools/math/LargeInt1.pri:173-184 (hash used by node_hash, Gap2Seq.cpp:837-843)
static inline uint64_t oahash64(uint64_t e) {
  uint64_t c = e;
  c = c ^ (c >> 14);
  c = (~c) + (c << 18);
  c = c ^ (c >> 31);
  c = c * 21;
  c = c ^ (c >> 11);
  c = c + (c << 6);
  c = c ^ (c >> 22);
  return c;
}

uint64_t Graph::revcomp(uint64_t x) const { return revcomp64(x, k_); }

void Graph::build_from_kmers(int k, const uint64_t *canon, size_t n) {
  k_ = k;
  mask_ = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
  size_t cap = 16;
  while (cap < 2 * n + 2) cap <<= 1;
  cap_mask_ = cap - 1;
  slots_.assign(cap, EMPTY);
  n_ = 0;
  for (size_t i = 0; i < n; i++) {
    uint64_t h = mix64(canon[i]) & cap_mask_;
    while (slots_[h] != EMPTY && slots_[h] != canon[i]) h = (h + 1) & cap_mask_;
    if (slots_[h] == EMPTY) { slots_[h] = canon[i]; n_++; }
  }
}

void Graph::build_from_reads(int k, int solid, const std::vector<std::string> &reads) {
  const uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
  std::vector<uint64_t> all;
  size_t total = 0;
  for (const std::string &r : reads) if ((int)r.size() >= k) total += r.size() - k + 1;
  all.reserve(total);
  for (const std::string &r : reads) {
    uint64_t fwd = 0, rev = 0;
    int run = 0;
    for (size_t i = 0; i < r.size(); i++) {
      unsigned char c = (unsigned char)r[i];
      if ((c >> 3) & 1) { run = 0; fwd = rev = 0; continue; }       // Data.hpp:176 invalid flag
      uint64_t nt = (c >> 1) & 3;
      fwd = ((fwd << 2) | nt) & mask;
      rev = (rev >> 2) | ((nt ^ 2) << (2 * (k - 1)));
      if (++run >= k) all.push_back(fwd < rev ? fwd : rev);
    }
  }
  std::sort(all.begin(), all.end());
  std::vector<uint64_t> keep;
  for (size_t i = 0; i < all.size();) {
    size_t j = i;
    while (j < all.size() && all[j] == all[i]) j++;
    if ((long long)(j - i) >= solid) keep.push_back(all[i]);
    i = j;
  }
  std::vector<uint64_t>().swap(all);
  build_from_kmers(k, keep.data(), keep.size());
}

bool Graph::contains(uint64_t canon) const {
  if (slots_.empty()) return false;
  uint64_t h = mix64(canon) & cap_mask_;
  while (slots_[h] != EMPTY) {
    if (slots_[h] == canon) return true;
    h = (h + 1) & cap_mask_;
  }
  return false;
}

Node Graph::build_node(const char *s) const {
  uint64_t fwd = 0;
  for (int i = 0; i < k_; i++) fwd = (fwd << 2) | (((unsigned char)s[i] >> 1) & 3);
  uint64_t rev = revcomp64(fwd, k_);
  Node n;
  n.kmer = fwd < rev ? fwd : rev;         // ModelCanonical: value = min(forward, revcomp)
  n.strand = (fwd == n.kmer) ? 0 : 1;     // Graph.cpp:1742
  return n;
}

std::string Graph::to_string(const Node &n) const {
  static const char bin2nt[4] = {'A', 'C', 'T', 'G'};
  uint64_t v = n.strand == 0 ? n.kmer : revcomp64(n.kmer, k_);
  std::string s(k_, 'A');
  for (int i = k_ - 1; i >= 0; i--) { s[i] = bin2nt[v & 3]; v >>= 2; }
  return s;
}

int Graph::successors(const Node &n, Node out[4]) const {
  uint64_t graine = n.strand == 0 ? n.kmer : revcomp64(n.kmer, k_);
  int cnt = 0;
  for (uint64_t nt = 0; nt < 4; nt++) {
    uint64_t fwd = ((graine << 2) + nt) & mask_;
    uint64_t rev = revcomp64(fwd, k_);
    if (fwd < rev) { if (contains(fwd)) { out[cnt].kmer = fwd; out[cnt].strand = 0; cnt++; } }
    else           { if (contains(rev)) { out[cnt].kmer = rev; out[cnt].strand = 1; cnt++; } }
  }
  return cnt;
}

int Graph::predecessors(const Node &n, Node out[4]) const {
  uint64_t graine = n.strand == 0 ? n.kmer : revcomp64(n.kmer, k_);
  int cnt = 0;
  for (uint64_t nt = 0; nt < 4; nt++) {
    uint64_t fwd = ((graine >> 2) + (nt << ((k_ - 1) * 2))) & mask_;
    uint64_t rev = revcomp64(fwd, k_);
    if (fwd < rev) { if (contains(fwd)) { out[cnt].kmer = fwd; out[cnt].strand = 0; cnt++; } }
    else           { if (contains(rev)) { out[cnt].kmer = rev; out[cnt].strand = 1; cnt++; } }
  }
  return cnt;
}

uint8_t Graph::adjacency(uint64_t canon) const {
  uint8_t v = 0;
  for (uint64_t nt = 0; nt < 4; nt++) {
    uint64_t f = ((canon << 2) + nt) & mask_;
    uint64_t r = revcomp64(f, k_);
    if (contains(f < r ? f : r)) v |= (uint8_t)(1u << nt);
    f = ((canon >> 2) + (nt << ((k_ - 1) * 2))) & mask_;
    r = revcomp64(f, k_);
    if (contains(f < r ? f : r)) v |= (uint8_t)(1u << (4 + nt));
  }
  return v;
}

void Graph::sorted_kmers(std::vector<uint64_t> &out) const {
  out.clear();
  for (uint64_t s : slots_) if (s != EMPTY) out.push_back(s);
  std::sort(out.begin(), out.end());
}

/* --------------------------------------------------------------- DP rows */

// Gap2Seq.cpp:488-659 (map_element): sparse (level,count) rows, one per strand.
struct RowF {
  std::vector<std::pair<int, int>> s[2];
  int get(int strand, int i) const {
    const std::vector<std::pair<int, int>> &v = s[strand];
    if (v.empty()) return 0;
    if (v.back().first == i) return v.back().second;
    auto it = std::lower_bound(v.begin(), v.end(), std::make_pair(i, -2147483647 - 1));
    return (it != v.end() && it->first == i) ? it->second : 0;
  }
  // append or overwrite the last entry; returns true when a new cell was created
  bool set(int strand, int i, int val) {
    std::vector<std::pair<int, int>> &v = s[strand];
    if (!v.empty() && v.back().first == i) { v.back().second = val; return false; }
    v.push_back(std::make_pair(i, val));
    return true;
  }
};

// Gap2Seq.cpp:663-834 (map_element2): existence only.
struct RowR {
  std::vector<int> s[2];
  bool set(int strand, int i) {
    std::vector<int> &v = s[strand];
    if (!v.empty() && v.back() == i) return false;
    v.push_back(i);
    return true;
  }
};

struct NodeHash { size_t operator()(const Node &n) const { return (size_t)oahash64(n.kmer); } };
struct NodeEq { bool operator()(const Node &a, const Node &b) const { return a.kmer == b.kmer; } };  // Graph.hpp:154
typedef std::unordered_set<Node, NodeHash, NodeEq> NodeSet;   // same hash/equality as Gap2Seq.cpp:872-873

// insert that also reports the strand collision the reference silently drops
static inline void border_insert(NodeSet &s, const Node &n, bool &collision) {
  std::pair<NodeSet::iterator, bool> r = s.insert(n);
  if (!r.second && r.first->strand != n.strand) collision = true;
}

/* ------------------------------------------------ subgraph (Boost stand-in) */

struct MiniGraph {
  struct E { size_t src, tgt, id; };
  std::vector<std::vector<E>> out, in;
  size_t nedges = 0, next_id = 0;
  size_t add_vertex() { out.emplace_back(); in.emplace_back(); return out.size() - 1; }
  void add_edge(size_t u, size_t v) { E e{u, v, next_id++}; out[u].push_back(e); in[v].push_back(e); nedges++; }
  bool has_edge(size_t u, size_t v) const { for (const E &e : out[u]) if (e.tgt == v) return true; return false; }
  static void erase_id(std::vector<E> &v, size_t id) { for (size_t i = 0; i < v.size(); i++) if (v[i].id == id) { v.erase(v.begin() + i); return; } }
  void clear_vertex(size_t v) {
    std::vector<E> o = out[v], i = in[v];
    for (const E &e : o) { erase_id(in[e.tgt], e.id); nedges--; }
    out[v].clear();
    for (const E &e : i) { if (e.src == v) continue; erase_id(out[e.src], e.id); nedges--; }
    in[v].clear();
  }
  void remove_edge(const E &e) { erase_id(out[e.src], e.id); erase_id(in[e.tgt], e.id); nedges--; }
  size_t nv() const { return out.size(); }
};

static size_t tarjan(const MiniGraph &g, std::vector<size_t> &comp) {
  const size_t n = g.nv(), NONE = (size_t)-1;
  comp.assign(n, 0);
  std::vector<size_t> index(n, NONE), low(n, 0), pos(n, 0), stack, call;
  std::vector<char> on(n, 0);
  size_t counter = 0, ncomp = 0;
  for (size_t r = 0; r < n; r++) {
    if (index[r] != NONE) continue;
    call.push_back(r); index[r] = low[r] = counter++; stack.push_back(r); on[r] = 1;
    while (!call.empty()) {
      size_t v = call.back();
      if (pos[v] < g.out[v].size()) {
        size_t w = g.out[v][pos[v]++].tgt;
        if (index[w] == NONE) { index[w] = low[w] = counter++; stack.push_back(w); on[w] = 1; call.push_back(w); }
        else if (on[w]) low[v] = std::min(low[v], index[w]);
      } else {
        call.pop_back();
        if (!call.empty()) low[call.back()] = std::min(low[call.back()], low[v]);
        if (low[v] == index[v]) {
          for (;;) { size_t w = stack.back(); stack.pop_back(); on[w] = 0; comp[w] = ncomp; if (w == v) break; }
          ncomp++;
        }
      }
    }
  }
  return ncomp;
}

// Kahn order; the branch counter of Gap2Seq.cpp:1412-1434 is independent of
// which topological order is used.
static void topo_order(const MiniGraph &g, std::vector<size_t> &order) {
  const size_t n = g.nv();
  std::vector<size_t> indeg(n);
  order.clear();
  std::vector<size_t> q;
  for (size_t v = 0; v < n; v++) { indeg[v] = g.in[v].size(); if (!indeg[v]) q.push_back(v); }
  for (size_t h = 0; h < q.size(); h++) {
    size_t v = q[h];
    order.push_back(v);
    for (const MiniGraph::E &e : g.out[v]) if (--indeg[e.tgt] == 0) q.push_back(e.tgt);
  }
}

/* -------------------------------------------------------------- fill_gap */

static inline int sat_add(int a, int b) { return (a + b > MAX_PATHS) ? MAX_PATHS : a + b; }  // Gap2Seq.cpp:1058

FillResult Oracle::fill_gap(const std::string &kmer_left, const std::string &kmer_right, int gap_len, int k,
                            int gap_err, int left_max_fuz, int right_max_fuz, bool skip_confident,
                            bool all_paths) const {
  FillResult res;
  memset(&res.substats, 0, sizeof(res.substats));
  res.count = 0; res.left_fuz = res.right_fuz = 0; res.substats_valid = false;
  res.cells_right = res.cells_left = res.nodes_right = res.nodes_left = 0;
  res.order_dependent = false; res.unable_to_backtrace = false;
  res.n_path_lengths = 0; res.path_lengths[0] = res.path_lengths[1] = -1; res.reached_fuz = 0;
  res.forward_count = 0; res.rand_calls = 0;

  const int right_half = right_max_fuz + (int)ceilf((gap_len + gap_err) / 2.f);   // Gap2Seq.cpp:862
  const int left_half = left_max_fuz + (int)floorf((gap_len + gap_err) / 2.f);    // Gap2Seq.cpp:863

  NodeSet border, nextBorder;
  std::unordered_map<uint64_t, RowR> rsRight;

  /* ---- Phase R: walks backwards from the staggered right starts (Gap2Seq.cpp:871-982) */
  Node node = g_.build_node(kmer_right.c_str() + (kmer_right.length() - k));
  int currentD = 0;
  if (g_.contains(node.kmer)) {
    border_insert(border, node, res.order_dependent);
    if (rsRight[node.kmer].set(node.strand, currentD)) res.cells_right++;
  }
  while (currentD < right_half) {
    currentD++;
    for (NodeSet::iterator it = border.begin(); it != border.end(); ++it) {
      Node nb[4];
      int nn = g_.predecessors(*it, nb);
      for (int i = 0; i < nn; i++) {
        if (rsRight[nb[i].kmer].set(nb[i].strand, currentD)) res.cells_right++;
        border_insert(nextBorder, nb[i], res.order_dependent);
      }
    }
    border.clear();
    border.swap(nextBorder);
    if (currentD <= right_max_fuz) {
      Node s = g_.build_node(kmer_right.c_str() + (kmer_right.length() - k - currentD));
      if (g_.contains(s.kmer)) {
        border_insert(border, s, res.order_dependent);
        if (rsRight[s.kmer].set(s.strand, currentD)) res.cells_right++;
      }
    }
  }
  res.nodes_right = rsRight.size();

  /* ---- Phase F: forward walk counts by exact length (Gap2Seq.cpp:984-1167) */
  border.clear();
  nextBorder.clear();
  int count = 0;
  std::unordered_map<uint64_t, RowF> rsLeft;
  node = g_.build_node(kmer_left.c_str());
  currentD = 0;
  if (g_.contains(node.kmer)) {
    border_insert(border, node, res.order_dependent);
    if (rsLeft[node.kmer].set(node.strand, currentD, 1)) res.cells_left++;
  }
  currentD++;

  Node reachedTarget; reachedTarget.kmer = 0; reachedTarget.strand = 0;
  int reachedFuz = 0;
  std::vector<int> pathLengths;
  const int prune_from = gap_len / 2 + gap_err / 2 + left_max_fuz;   // Gap2Seq.cpp:1050 (two integer divisions)

  while (currentD <= right_half + left_half) {
    for (NodeSet::iterator it = border.begin(); it != border.end(); ++it) {
      const Node n = *it;
      Node nb[4];
      int nn = g_.successors(n, nb);
      int num_paths = rsLeft[n.kmer].get(n.strand, currentD - 1);
      for (int i = 0; i < nn; i++) {
        if (currentD < prune_from || rsRight.find(nb[i].kmer) != rsRight.end()) {
          RowF &me = rsLeft[nb[i].kmer];
          if (me.set(nb[i].strand, currentD, sat_add(me.get(nb[i].strand, currentD), num_paths))) res.cells_left++;
          border_insert(nextBorder, nb[i], res.order_dependent);
        }
      }
    }
    border.clear();
    border.swap(nextBorder);

    if (currentD <= left_max_fuz) {   // next staggered left start, count OVERWRITTEN with 1 (Gap2Seq.cpp:1082-1105)
      Node s = g_.build_node(kmer_left.c_str() + currentD);
      if (g_.contains(s.kmer)) {
        border_insert(border, s, res.order_dependent);
        if (rsLeft[s.kmer].set(s.strand, currentD, 1)) res.cells_left++;
      }
    }

    if (pathLengths.empty() && currentD >= gap_len + left_max_fuz + right_max_fuz) {   // Gap2Seq.cpp:1108-1159
      int err = currentD - gap_len - (left_max_fuz + right_max_fuz);
      for (int j = 0; j <= right_max_fuz && count == 0; j++) {
        reachedTarget = g_.build_node(kmer_right.c_str() + j);
        std::unordered_map<uint64_t, RowF>::iterator f = rsLeft.find(reachedTarget.kmer);
        if (f == rsLeft.end()) continue;
        const RowF &right = f->second;
        int len1 = gap_len + left_max_fuz + j + err;
        int len2 = gap_len + left_max_fuz + j - err;
        reachedFuz = j;
        int c1 = right.get(reachedTarget.strand, len1);
        if (c1 >= 1) { count = sat_add(count, c1); pathLengths.push_back(len1); }
        if (len2 != len1 && len2 >= 0) {
          int c2 = right.get(reachedTarget.strand, len2);
          if (c2 >= 1) { count = sat_add(count, c2); pathLengths.push_back(len2); }
        }
      }
      if (!all_paths && !pathLengths.empty()) break;
    }
    currentD++;
  }
  res.nodes_left = rsLeft.size();
  res.forward_count = count;
  res.n_path_lengths = (int)pathLengths.size();
  for (size_t i = 0; i < pathLengths.size() && i < 2; i++) res.path_lengths[i] = pathLengths[i];
  res.reached_fuz = reachedFuz;

  if (count > 0 && !pathLengths.empty()) {
    res.right_fuz = reachedFuz;
    std::unordered_map<uint64_t, size_t> node2v;
    std::vector<int> branch;
    MiniGraph sub;

    if (!skip_confident) {
      /* ---- Phase B: union of accepted walks collapsed by canonical k-mer (Gap2Seq.cpp:1181-1312) */
      NodeSet backBorder, nextBackBorder;
      const size_t sink = sub.add_vertex();
      const size_t source = sub.add_vertex();
      int d = left_max_fuz + gap_len + gap_err + right_max_fuz;
      if (all_paths) count = 0;
      auto vertex_of = [&](const Node &n) -> size_t {
        std::unordered_map<uint64_t, size_t>::iterator it = node2v.find(n.kmer);
        if (it != node2v.end()) return it->second;
        size_t v = sub.add_vertex();
        node2v[n.kmer] = v;
        return v;
      };
      while (d >= 0) {
        if (all_paths) {
          if (d >= left_max_fuz + gap_len - gap_err) {
            for (int j = 0; j < right_max_fuz; j++) {     // strictly less: Gap2Seq.cpp:1197
              Node rn = g_.build_node(kmer_right.c_str() + j);
              if (j < right_max_fuz - 1) { if (g_.contains(rn.kmer)) continue; }   // Gap2Seq.cpp:1201-1206
              std::unordered_map<uint64_t, RowF>::iterator f = rsLeft.find(rn.kmer);
              if (f == rsLeft.end()) continue;
              int c = f->second.get(rn.strand, d);
              if (c >= 1) {
                count = sat_add(count, c);
                border_insert(backBorder, rn, res.order_dependent);
                size_t v = vertex_of(rn);
                if (!sub.has_edge(v, sink)) sub.add_edge(v, sink);
              }
            }
          }
        } else {
          for (size_t j = 0; j < pathLengths.size(); j++) {
            if (pathLengths[j] == d) {
              border_insert(backBorder, reachedTarget, res.order_dependent);
              size_t v = vertex_of(reachedTarget);
              if (!sub.has_edge(v, sink)) sub.add_edge(v, sink);
            }
          }
        }
        Node lnode; lnode.kmer = 0; lnode.strand = 0;   // Node_t() default: kmer value 0
        if (d <= left_max_fuz) lnode = g_.build_node(kmer_left.c_str() + d);
        for (NodeSet::iterator it = backBorder.begin(); it != backBorder.end(); ++it) {
          const Node cur = *it;
          if (d > left_max_fuz || cur.kmer != lnode.kmer) {
            Node nb[4];
            int nn = g_.predecessors(cur, nb);
            for (int i = 0; i < nn; i++) {
              std::unordered_map<uint64_t, RowF>::iterator f = rsLeft.find(nb[i].kmer);
              if (f == rsLeft.end()) continue;
              if (f->second.get(nb[i].strand, d - 1) > 0) {
                border_insert(nextBackBorder, nb[i], res.order_dependent);
                size_t u = vertex_of(nb[i]);
                size_t v = vertex_of(cur);
                if (!sub.has_edge(u, v)) sub.add_edge(u, v);
              }
            }
          } else {
            size_t v = vertex_of(cur);
            if (!sub.has_edge(source, v)) sub.add_edge(source, v);
          }
        }
        backBorder.clear();
        backBorder.swap(nextBackBorder);
        d--;
      }

      /* ---- SCC contraction, self loops, stats, branch counter (Gap2Seq.cpp:1314-1434) */
      std::vector<size_t> comp;
      const size_t ncomp = tarjan(sub, comp);
      std::vector<int> csize(ncomp, 0);
      const size_t nreal = sub.nv();
      size_t nreal_edges = sub.nedges, n_nontriv = 0, size_nontriv = 0;
      if (ncomp != nreal) {
        for (size_t i = 0; i < nreal; i++) csize[comp[i]]++;
        std::vector<size_t> cnode(ncomp, 0);
        for (size_t i = 0; i < ncomp; i++) if (csize[i] > 1) { cnode[i] = sub.add_vertex(); n_nontriv++; }
        for (size_t i = 0; i < nreal; i++) {
          if (csize[comp[i]] <= 1) continue;
          size_nontriv++;
          std::vector<MiniGraph::E> ins = sub.in[i];   // copies: adding edges elsewhere never touches i's own lists
          for (const MiniGraph::E &e : ins) {
            size_t s = e.src;
            if (comp[i] == comp[s]) continue;
            if (csize[comp[s]] > 1) { if (s < i) sub.add_edge(cnode[comp[s]], cnode[comp[i]]); }
            else sub.add_edge(s, cnode[comp[i]]);
          }
          std::vector<MiniGraph::E> outs = sub.out[i];
          for (const MiniGraph::E &e : outs) {
            size_t t = e.tgt;
            if (comp[i] == comp[t]) continue;
            if (csize[comp[t]] > 1) { if (t < i) sub.add_edge(cnode[comp[i]], cnode[comp[t]]); }
            else sub.add_edge(cnode[comp[i]], t);
          }
        }
        for (size_t i = 0; i < nreal; i++) if (csize[comp[i]] > 1) sub.clear_vertex(i);
      } else {
        for (size_t i = 0; i < ncomp; i++) csize[i] = 1;
      }
      for (size_t i = 0; i < nreal; i++) {
        if (csize[comp[i]] > 1) continue;
        std::vector<MiniGraph::E> loops;
        for (const MiniGraph::E &e : sub.out[i]) if (e.src == e.tgt) loops.push_back(e);
        for (const MiniGraph::E &e : loops) sub.remove_edge(e);
        nreal_edges -= loops.size();
      }
      res.substats.vertices = nreal;
      res.substats.edges = nreal_edges;
      res.substats.nontrivial_components = n_nontriv;
      res.substats.size_nontrivial_components = size_nontriv;
      res.substats.vertices_final = sub.nv() - size_nontriv;
      res.substats.edges_final = sub.nedges;
      res.substats_valid = true;

      branch.assign(sub.nv(), 0);
      std::vector<size_t> order;
      topo_order(sub, order);
      int branchcount = 1;
      for (size_t v : order) {
        size_t di = sub.in[v].size(), dout = sub.out[v].size();
        if (di >= 1 || dout >= 1) {
          if (di > 1) branchcount -= (int)di - 1;
          branch[v] = branchcount;
          if (dout > 1) branchcount += (int)dout - 1;
        }
      }
    }

    /* ---- Phase T: random traceback and case mask (Gap2Seq.cpp:1437-1517) */
    res.rand_calls++;
    int d = pathLengths[rand() % pathLengths.size()];
    int lastSolid = d;
    Node current = reachedTarget;
    std::vector<char> fill((size_t)d + 1, '?');
    fill[d] = '\0';
    const int d0 = d;
    int left_fuz = 0;
    while (d >= 0) {
      std::string str = g_.to_string(current);
      if (d <= left_max_fuz) {
        Node lnode = g_.build_node(kmer_left.c_str() + d);
        if (lnode.kmer == current.kmer) { left_fuz = left_max_fuz - d; break; }
      }
      if (d > 0) {
        bool safe = skip_confident;
        if (!safe) {
          // node2boost[current] default-inserts vertex 0 (= sink) for unknown nodes
          std::unordered_map<uint64_t, size_t>::iterator it = node2v.find(current.kmer);
          size_t v = (it == node2v.end()) ? 0 : it->second;
          safe = branch[v] == 1;
        }
        if (safe) lastSolid = d;
        char c = str[str.length() - 1];
        fill[d - 1] = (d > lastSolid - k) ? (char)toupper(c) : (char)tolower(c);
        Node nb[4], back[4];
        int nn = g_.predecessors(current, nb), nback = 0;
        for (int i = 0; i < nn; i++) {
          std::unordered_map<uint64_t, RowF>::iterator f = rsLeft.find(nb[i].kmer);
          if (f != rsLeft.end() && f->second.get(nb[i].strand, d - 1) > 0) back[nback++] = nb[i];
        }
        if (nback == 0) {   // "Unable to backtrace!" (Gap2Seq.cpp:1493-1510)
          res.unable_to_backtrace = true;
          res.count = 0;
          for (auto &kv : rsLeft) res.cells_left += 0 * kv.second.s[0].size();
          return res;
        }
        res.rand_calls++;
        current = back[rand() % nback];
      }
      d--;
    }
    res.left_fuz = left_fuz;
    res.fill.assign(&fill[left_max_fuz - left_fuz], &fill[d0]);
  }
  res.count = count;
  return res;
}

/* --------------------------------------------- statistics line + scanner */

// Gap2Seq.cpp:100-156
static void format_statistics(std::string &log, int filledStart, int gapStart, int gapEnd, int paths,
                              const std::string &fillstr, int k, int lmf, int rmf, int lf, int rf, bool skip_confident,
                              bool unique_paths, const SubgraphStats &ss, int gap, const std::string &comment) {
  char buf[1024];
  if (paths > 0 && (!unique_paths || paths == 1)) {
    int filledLen = (int)fillstr.size() - k;
    if (!skip_confident) {
      int upper = 0, lower = 0;
      for (int j = 0; j < filledLen; j++) { if (isupper((unsigned char)fillstr[j])) upper++; else lower++; }
      snprintf(buf, sizeof buf,
               " GapStart: %d GapEnd: %d GapLength: %d PathsFound: %d FilledStart: %d FilledEnd: %d FilledGapLength: %d"
               " LeftMaxFuz: %d LeftFuz: %d RightMaxFuz: %d RightFuz: %d ConfidentBases: %d TotalBases: %d\n",
               gapStart, gapEnd, gap, paths, filledStart, filledStart + filledLen, filledLen, lmf, lf, rmf, rf, upper,
               upper + lower);
      log += "Scaffold: " + comment + buf;
      snprintf(buf, sizeof buf,
               "SubgraphStats: Vertices: %llu Edges: %llu NontrivialStrongComponents: %llu"
               " SizeNontrivialStrongComponents: %llu VerticesFinal: %llu EdgesFinal: %llu\n",
               (unsigned long long)ss.vertices, (unsigned long long)ss.edges,
               (unsigned long long)ss.nontrivial_components, (unsigned long long)ss.size_nontrivial_components,
               (unsigned long long)ss.vertices_final, (unsigned long long)ss.edges_final);
      log += buf;
    } else {
      snprintf(buf, sizeof buf,
               " GapStart: %d GapEnd: %d GapLength: %d PathsFound: %d FilledStart: %d FilledEnd: %d FilledGapLength: %d"
               " LeftFuz: %d RightFuz: %d\n",
               gapStart, gapEnd, gap, paths, filledStart, filledStart + filledLen, filledLen, lf, rf);
      log += "Scaffold: " + comment + buf;
    }
  } else {
    snprintf(buf, sizeof buf,
             " GapStart: %d GapEnd: %d GapLength: %d PathsFound: 0 FilledStart: 0 FilledEnd: 0 FilledGapLength: 0"
             " LeftMaxFuz: %d LeftFuz: %d RightMaxFuz: %d RightFuz: %d%s\n",
             gapStart, gapEnd, gap, lmf, lf, rmf, rf, paths == -1 ? " Memory limit exceeded" : "");
    log += "Scaffold: " + comment + buf;
  }
}

// Gap2Seq.cpp:329-432
std::string fill_record(const Oracle &o, const std::string &seq, const std::string &comment, const RunOptions &opt,
                        std::string &log, int &gapcount, int &filledgapcount, std::vector<FillResult> *details) {
  const int k = opt.k;
  std::string filledSeq;
  int prevGapEnd = 0;
  size_t i = 0;
  while (i < seq.size()) {
    if (seq[i] == 'N' || seq[i] == 'n') {
      gapcount++;
      int lmf = std::min(((int)i + k) - prevGapEnd, opt.max_fuz);
      int kmer_start = (int)i - k - lmf;
      int gap = 0;
      while (i < seq.size() && (seq[i] == 'N' || seq[i] == 'n')) { i++; gap++; }
      int rmf = std::min((int)seq.size() - ((int)i + k), opt.max_fuz);
      bool ok = i + k + rmf <= seq.size();
      for (int j = 0; j < k + rmf && ok; j++) if (seq[i + j] == 'N' || seq[i + j] == 'n') ok = false;
      if (kmer_start >= prevGapEnd && ok) {
        FillResult r = o.fill_gap(seq.substr(kmer_start, k + lmf), seq.substr(i, k + rmf), gap, k, opt.d_err, lmf, rmf,
                                  opt.skip_confident, opt.all_paths);
        if (details) details->push_back(r);
        if (opt.stats) {   // the -DSTATS line is printed inside fill_gap, i.e. before the Scaffold: line (Gap2Seq.cpp:1548-1553)
          char buf[160];
          snprintf(buf, sizeof buf, "Nodes: %llu Nonzeros: %llu\n",
                   (unsigned long long)(r.nodes_left + r.nodes_right), (unsigned long long)(r.cells_left + r.cells_right));
          log += buf;
        }
        int filledStart = (int)filledSeq.length() + kmer_start + k + lmf - r.left_fuz - prevGapEnd;
        int gapStart = kmer_start + k + lmf;
        format_statistics(log, filledStart, gapStart, (int)i, r.count, r.fill, k, lmf, rmf, r.left_fuz, r.right_fuz,
                          opt.skip_confident, opt.unique_paths, r.substats, gap, comment);
        if (r.count > 0 && (!opt.unique_paths || r.count == 1)) {
          filledgapcount++;
          // '=' not '+=': the prefix of a multi-gap record is dropped (Gap2Seq.cpp:396)
          filledSeq = seq.substr(prevGapEnd, kmer_start + k + lmf - r.left_fuz - prevGapEnd) + r.fill;
          filledSeq = filledSeq.substr(0, filledSeq.length() - k);
          i += r.right_fuz;
        } else {
          filledSeq = filledSeq + seq.substr(prevGapEnd, kmer_start + k + lmf + gap - prevGapEnd);
        }
      } else {
        filledSeq = filledSeq + seq.substr(prevGapEnd, kmer_start + k + lmf + gap - prevGapEnd);
      }
      prevGapEnd = (int)i;
    } else {
      i++;
    }
  }
  filledSeq = filledSeq + seq.substr(prevGapEnd, seq.length() - prevGapEnd);
  return filledSeq;
}

}  // namespace g2s_oracle
