// Minimal stand-in for the part of Boost.Graph that the reference's
// src/Gap2Seq.cpp uses (includes at Gap2Seq.cpp:29-31, typedefs at
// Gap2Seq.hpp:33-35).  Boost.Graph is neither vendored by the reference nor
// installed in this image, so the reference binary under oracle/_ref/ is
// compiled against this header instead.  TEST INFRASTRUCTURE ONLY: nothing in
// the product links or includes it.
//
// What fill_gap derives from the graph library (component sizes, the
// "branch" counter along a topological order, vertex/edge counts) does not
// depend on the vertex visiting order of the DFS, so
// a straightforward Tarjan + DFS topological sort is sufficient.
#ifndef G2S_ORACLE_BOOST_SHIM_ADJACENCY_LIST_HPP
#define G2S_ORACLE_BOOST_SHIM_ADJACENCY_LIST_HPP

#include <cstddef>
#include <utility>
#include <vector>
#include <algorithm>

namespace boost {

struct vecS {};
struct bidirectionalS {};
enum vertex_index_t { vertex_index };

struct shim_edge {
  std::size_t src, tgt, id;
  bool operator==(const shim_edge &o) const { return id == o.id; }
  bool operator!=(const shim_edge &o) const { return id != o.id; }
};

template <class OutEdgeList, class VertexList, class Directed>
class adjacency_list {
 public:
  typedef std::size_t vertex_descriptor;
  typedef shim_edge edge_descriptor;
  typedef std::vector<shim_edge>::const_iterator out_edge_iterator;
  typedef std::vector<shim_edge>::const_iterator in_edge_iterator;

  explicit adjacency_list(std::size_t n = 0) : out_(n), in_(n), nedges_(0), next_id_(0) {}

  std::vector<std::vector<shim_edge> > out_, in_;
  std::size_t nedges_, next_id_;
};

template <class G>
struct graph_traits {
  typedef typename G::vertex_descriptor vertex_descriptor;
  typedef typename G::edge_descriptor edge_descriptor;
  typedef typename G::out_edge_iterator out_edge_iterator;
  typedef typename G::in_edge_iterator in_edge_iterator;
};

#define G2S_SHIM_TPL template <class A, class B, class C>
#define G2S_SHIM_G adjacency_list<A, B, C>

G2S_SHIM_TPL std::size_t add_vertex(G2S_SHIM_G &g) {
  g.out_.push_back(std::vector<shim_edge>());
  g.in_.push_back(std::vector<shim_edge>());
  return g.out_.size() - 1;
}

G2S_SHIM_TPL std::pair<shim_edge, bool> add_edge(std::size_t u, std::size_t v, G2S_SHIM_G &g) {
  shim_edge e = {u, v, g.next_id_++};
  g.out_[u].push_back(e);
  g.in_[v].push_back(e);
  g.nedges_++;
  return std::make_pair(e, true);
}

G2S_SHIM_TPL std::pair<shim_edge, bool> edge(std::size_t u, std::size_t v, const G2S_SHIM_G &g) {
  for (std::size_t i = 0; i < g.out_[u].size(); i++)
    if (g.out_[u][i].tgt == v) return std::make_pair(g.out_[u][i], true);
  shim_edge none = {0, 0, (std::size_t)-1};
  return std::make_pair(none, false);
}

G2S_SHIM_TPL std::size_t num_vertices(const G2S_SHIM_G &g) { return g.out_.size(); }
G2S_SHIM_TPL std::size_t num_edges(const G2S_SHIM_G &g) { return g.nedges_; }
G2S_SHIM_TPL std::size_t in_degree(std::size_t v, const G2S_SHIM_G &g) { return g.in_[v].size(); }
G2S_SHIM_TPL std::size_t out_degree(std::size_t v, const G2S_SHIM_G &g) { return g.out_[v].size(); }

G2S_SHIM_TPL std::pair<std::vector<shim_edge>::const_iterator, std::vector<shim_edge>::const_iterator>
out_edges(std::size_t v, const G2S_SHIM_G &g) {
  return std::make_pair(g.out_[v].begin(), g.out_[v].end());
}
G2S_SHIM_TPL std::pair<std::vector<shim_edge>::const_iterator, std::vector<shim_edge>::const_iterator>
in_edges(std::size_t v, const G2S_SHIM_G &g) {
  return std::make_pair(g.in_[v].begin(), g.in_[v].end());
}

G2S_SHIM_TPL std::size_t source(const shim_edge &e, const G2S_SHIM_G &) { return e.src; }
G2S_SHIM_TPL std::size_t target(const shim_edge &e, const G2S_SHIM_G &) { return e.tgt; }

namespace shim_detail {
inline void erase_id(std::vector<shim_edge> &v, std::size_t id) {
  for (std::size_t i = 0; i < v.size(); i++)
    if (v[i].id == id) { v.erase(v.begin() + i); return; }
}
}  // namespace shim_detail

G2S_SHIM_TPL void remove_edge(const shim_edge &e, G2S_SHIM_G &g) {
  shim_detail::erase_id(g.out_[e.src], e.id);
  shim_detail::erase_id(g.in_[e.tgt], e.id);
  g.nedges_--;
}

G2S_SHIM_TPL void clear_vertex(std::size_t v, G2S_SHIM_G &g) {
  std::vector<shim_edge> o = g.out_[v], i = g.in_[v];
  for (std::size_t j = 0; j < o.size(); j++) {
    shim_detail::erase_id(g.in_[o[j].tgt], o[j].id);
    g.nedges_--;
  }
  g.out_[v].clear();
  for (std::size_t j = 0; j < i.size(); j++) {
    if (i[j].src == v) continue;  // self loop, already counted through out_
    shim_detail::erase_id(g.out_[i[j].src], i[j].id);
    g.nedges_--;
  }
  g.in_[v].clear();
}

// property-map plumbing used only to hand a std::vector to strong_components
struct shim_identity_map {};
G2S_SHIM_TPL shim_identity_map get(vertex_index_t, const G2S_SHIM_G &) { return shim_identity_map(); }

template <class It>
struct shim_iter_pmap {
  It it;
};
template <class It>
shim_iter_pmap<It> make_iterator_property_map(It it, shim_identity_map) {
  shim_iter_pmap<It> p = {it};
  return p;
}

// boost::tie for the std::pair<iterator, iterator> returned by in/out_edges
template <class T1, class T2>
struct shim_tier {
  T1 &a;
  T2 &b;
  shim_tier(T1 &a_, T2 &b_) : a(a_), b(b_) {}
  template <class U1, class U2>
  shim_tier &operator=(const std::pair<U1, U2> &p) {
    a = p.first;
    b = p.second;
    return *this;
  }
};
template <class T1, class T2>
shim_tier<T1, T2> tie(T1 &a, T2 &b) { return shim_tier<T1, T2>(a, b); }

#undef G2S_SHIM_TPL
#undef G2S_SHIM_G

}  // namespace boost
#endif
