// Stand-in for boost/graph/topological_sort.hpp (see adjacency_list.hpp).
// Like Boost's, it emits vertices in REVERSE topological order (DFS finish
// order); src/Gap2Seq.cpp:1418-1421 walks the result backwards.
#ifndef G2S_ORACLE_BOOST_SHIM_TOPOLOGICAL_SORT_HPP
#define G2S_ORACLE_BOOST_SHIM_TOPOLOGICAL_SORT_HPP
#include <boost/graph/adjacency_list.hpp>

namespace boost {

template <class A, class B, class C, class OutIt>
void topological_sort(const adjacency_list<A, B, C> &g, OutIt out) {
  const std::size_t n = g.out_.size();
  std::vector<char> seen(n, 0);
  std::vector<std::size_t> pos(n, 0), call;
  for (std::size_t r = 0; r < n; r++) {
    if (seen[r]) continue;
    seen[r] = 1;
    call.push_back(r);
    while (!call.empty()) {
      std::size_t v = call.back();
      if (pos[v] < g.out_[v].size()) {
        std::size_t w = g.out_[v][pos[v]++].tgt;
        if (!seen[w]) {
          seen[w] = 1;
          call.push_back(w);
        }
      } else {
        call.pop_back();
        *out++ = v;
      }
    }
  }
}

}  // namespace boost
#endif
