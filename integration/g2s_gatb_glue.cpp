// g2s_gatb_glue.cpp - INTEGRATION.md form 2, compiled INTO the reference binary next to its unchanged Gap2Seq.cpp
// (oracle/build_ref.sh builds oracle/_ref/bin/Gap2Seq-core-g2s that way; this file is the only new source).
//
//   * first call: the solid k-mers of the caller's GATB graph (graph.iterator(), Graph.cpp:2243-2362) are imported into
//     the device table once (g2s_import_kmers);
//   * every call: the caller's graph answers contains() for the flank k-mers fill_gap looks up (src/Gap2Seq.cpp:884,
//     957, 1000, 1086 - Bloom AND NOT cFP AND MPHF, Graph.hpp:1249-1272) and the answers travel with the gap
//     (g2s_gap.present_bits): where the Bloom filter reports a k-mer the exact solid set does not hold, the engine
//     flags the gap G2S_ST_PHANTOM_FLANK and this glue says so on stderr - reported, never silent;
//   * the call itself is g2s_fill_gap_ex: same buffer contract, same return value, same libc rand() consumption.
// The reference calls fill_gap from -nb-cores Dispatcher threads; the engine is one device queue, calls are serialised
// here (parity needs -nb-cores 1 anyway).
#include "g2s_gatb_glue.h"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <vector>

#include "Gap2Seq.hpp"   // struct subgraph_stats (src/Gap2Seq.hpp:51-59), the reference's own header
#include "g2s.h"

static g2s_engine *g_engine = NULL;
static std::mutex g_mu;

static g2s_engine *engine_for(const Graph &graph, int k) {
  if (g_engine) return g_engine;
  g2s_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = G2S_ABI_VERSION;
  cfg.device = getenv("G2S_DEVICE") ? atoi(getenv("G2S_DEVICE")) : 0;
  cfg.k = k;
  cfg.max_fuz = 31;
  cfg.host_threads = 1;
  if (g2s_create(&cfg, &g_engine) != G2S_OK) { std::cerr << "g2s: " << g2s_last_error() << std::endl; exit(EXIT_FAILURE); }
  std::vector<uint64_t> canon;
  GraphIterator<Node> it = graph.iterator();
  for (it.first(); !it.isDone(); it.next()) canon.push_back(it.item().kmer.get<LargeInt<1> >().getVal());   // k <= 31
  if (g2s_import_kmers(g_engine, canon.data(), canon.size()) != G2S_OK) {
    std::cerr << "g2s: " << g2s_last_error() << std::endl;
    exit(EXIT_FAILURE);
  }
  std::cerr << "g2s: imported " << canon.size() << " solid k-mers from the GATB graph" << std::endl;
  return g_engine;
}

int g2s_fill_gap_graph(const Graph &graph, const std::string &kmer_left, const std::string &kmer_right, int gap_len, int k,
                       int gap_err, int left_max_fuz, int right_max_fuz, int *left_fuz, int *right_fuz, long long max_mem,
                       char *fill, bool skip_confident, bool all_paths, struct subgraph_stats *substats) {
  std::lock_guard<std::mutex> lk(g_mu);
  g2s_engine *eng = engine_for(graph, k);
  uint64_t bits = 0;
  for (int d = 0; d <= left_max_fuz && d < 32; d++)
    if (graph.contains(graph.buildNode(kmer_left.substr(d, k).c_str()))) bits |= 1ULL << d;
  for (int j = 0; j <= right_max_fuz && j < 32; j++)
    if (graph.contains(graph.buildNode(kmer_right.substr(j, k).c_str()))) bits |= 1ULL << (32 + j);
  g2s_subgraph_stats ss;
  memset(&ss, 0, sizeof ss);
  int32_t status = 0;
  fill[0] = 0;
  const int s = g2s_fill_gap_ex(eng, kmer_left.c_str(), kmer_right.c_str(), gap_len, k, gap_err, left_max_fuz, right_max_fuz,
                                left_fuz, right_fuz, max_mem, fill, skip_confident ? 1 : 0, all_paths ? 1 : 0, &ss, bits, 1, &status);
  if (s < -1) { std::cerr << "g2s: " << g2s_last_error() << std::endl; exit(EXIT_FAILURE); }
  substats->vertices = ss.vertices; substats->edges = ss.edges;
  substats->nontrivial_components = ss.nontrivial_components;
  substats->size_nontrivial_components = ss.size_nontrivial_components;
  substats->vertices_final = ss.vertices_final; substats->edges_final = ss.edges_final;
  if (status & G2S_ST_PHANTOM_FLANK)
    std::cerr << "g2s: PHANTOM_FLANK: GATB's contains() and the solid k-mer set disagree on a flank k-mer of the gap after "
              << kmer_left.substr(0, 12) << "... (Bloom false positive, SURVEY Q13); filled on the exact set" << std::endl;
  if (status & G2S_ST_ORDER_DEPENDENT)
    std::cerr << "g2s: ORDER_DEPENDENT: both strands of a k-mer on one level (SURVEY Q3)" << std::endl;
  return s;
}
