// Stand-in for boost/graph/strong_components.hpp (see adjacency_list.hpp in
// this directory for why it exists).  Iterative Tarjan; component numbering is
// arbitrary, which is all src/Gap2Seq.cpp:1314-1383 relies on.
#ifndef G2S_ORACLE_BOOST_SHIM_STRONG_COMPONENTS_HPP
#define G2S_ORACLE_BOOST_SHIM_STRONG_COMPONENTS_HPP
#include <boost/graph/adjacency_list.hpp>

namespace boost {

template <class A, class B, class C, class It>
std::size_t strong_components(const adjacency_list<A, B, C> &g, shim_iter_pmap<It> comp) {
  const std::size_t n = g.out_.size(), NONE = (std::size_t)-1;
  std::vector<std::size_t> index(n, NONE), low(n, 0), pos(n, 0), stack, call;
  std::vector<char> on(n, 0);
  std::size_t counter = 0, ncomp = 0;
  for (std::size_t r = 0; r < n; r++) {
    if (index[r] != NONE) continue;
    call.push_back(r);
    index[r] = low[r] = counter++;
    stack.push_back(r);
    on[r] = 1;
    while (!call.empty()) {
      std::size_t v = call.back();
      if (pos[v] < g.out_[v].size()) {
        std::size_t w = g.out_[v][pos[v]++].tgt;
        if (index[w] == NONE) {
          index[w] = low[w] = counter++;
          stack.push_back(w);
          on[w] = 1;
          call.push_back(w);
        } else if (on[w]) {
          low[v] = std::min(low[v], index[w]);
        }
      } else {
        call.pop_back();
        if (!call.empty()) low[call.back()] = std::min(low[call.back()], low[v]);
        if (low[v] == index[v]) {
          for (;;) {
            std::size_t w = stack.back();
            stack.pop_back();
            on[w] = 0;
            *(comp.it + w) = ncomp;
            if (w == v) break;
          }
          ncomp++;
        }
      }
    }
  }
  return ncomp;
}

}  // namespace boost
#endif
