// g2s_oracle - CPU restatement of the reference hot path (Gap2Seq::fill_gap and
// the pieces of GATB it queries).  TEST INFRASTRUCTURE ONLY: it is the parity
// checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
// The product (gap2seq_b200/) never includes, links or executes this file.
//
// Parity status: PINNED against the unmodified reference binary built by
// oracle/build_ref.sh (oracle/_ref/bin/Gap2Seq-core, -nb-cores 1, fixed
// -randseed) on the synthetic workloads of tests/; see tests/test_oracle_vs_ref.py
// and the committed goldens under tests/golden/.  The reference itself ships no
// tests or golden vectors for this path.
//
// Reference locations restated here (paths relative to /root/reference):
//   src/Gap2Seq.cpp:858-1556   fill_gap            -> Oracle::fill_gap
//   src/Gap2Seq.cpp:488-834    map_element(2)      -> RowF / RowR
//   src/Gap2Seq.cpp:837-843    node_hash           -> NodeHash (oahash64)
//   src/Gap2Seq.cpp:285-438    execute() scanner   -> fill_record
//   src/Gap2Seq.cpp:100-156    print_statistics    -> format_statistics
//   GATB debruijn/impl/Graph.cpp:1284-1345 successors/predecessors (Bloom mode)
//   GATB debruijn/impl/Graph.cpp:1726-1744 buildNode; :2400-2414 toString
//   GATB tools/math/LargeInt1.pri:119-184 toString/revcomp64/oahash64
//   GATB tools/misc/api/Data.hpp:176      ASCII -> 2 bit ((c>>1)&3, invalid (c>>3)&1)
//   GATB kmer/impl/Sequence2SuperKmer.hpp:93 invalid k-mers are not counted
//
// Deliberate difference: graph membership is the EXACT solid k-mer set
// (abundance >= -solid).  GATB's contains() can return a Bloom false positive
// for a non-solid k-mer that is not a neighbour of a solid one;
// only flank k-mers are ever queried that way.
#ifndef G2S_ORACLE_HPP
#define G2S_ORACLE_HPP

#include <cstdint>
#include <string>
#include <vector>

namespace g2s_oracle {

static const int MAX_PATHS = (2147483647 / 2 - 1);  // Gap2Seq.cpp:38

struct Node {
  uint64_t kmer;  // canonical value (min of forward / reverse complement)
  int strand;     // 0 = STRAND_FORWARD, 1 = STRAND_REVCOMP
};

// Exact solid k-mer set with GATB's neighbour semantics.
class Graph {
 public:
  Graph() : k_(0), mask_(0), cap_mask_(0) {}
  // canonical k-mers in GATB 2-bit code (A=0,C=1,T=2,G=3)
  void build_from_kmers(int k, const uint64_t *canon, size_t n);
  // count canonical k-mers over reads, keep abundance >= solid (DSK "sum" solidity)
  void build_from_reads(int k, int solid, const std::vector<std::string> &reads);
  int k() const { return k_; }
  size_t size() const { return n_; }
  bool contains(uint64_t canon) const;
  Node build_node(const char *s) const;               // Graph.cpp:1726-1744
  std::string to_string(const Node &n) const;         // Graph.cpp:2400-2414
  int successors(const Node &n, Node out[4]) const;   // Graph.cpp:1284-1310
  int predecessors(const Node &n, Node out[4]) const; // Graph.cpp:1313-1345
  uint64_t revcomp(uint64_t x) const;                 // LargeInt1.pri:136-154
  void sorted_kmers(std::vector<uint64_t> &out) const;
  // 8-bit adjacency byte as GATB's precomputeAdjacency defines it
  // (Graph.cpp:3486-3530): low nibble = outgoing nt A,C,T,G of the canonical
  // orientation, high nibble = incoming.
  uint8_t adjacency(uint64_t canon) const;

 private:
  int k_;
  uint64_t mask_;
  uint64_t cap_mask_;
  size_t n_ = 0;
  std::vector<uint64_t> slots_;
};

struct SubgraphStats {  // Gap2Seq.hpp:51-59
  uint64_t vertices, edges, nontrivial_components, size_nontrivial_components, vertices_final, edges_final;
};

struct FillResult {
  int count;           // return value of fill_gap
  int left_fuz, right_fuz;
  std::string fill;    // == std::string(&fill[left_max_fuz - left_fuz]) (ends with the right k-mer); empty when not traced
  SubgraphStats substats;
  bool substats_valid;
  uint64_t cells_right, cells_left;   // -DSTATS "Nonzeros" split by table
  uint64_t nodes_right, nodes_left;   // -DSTATS "Nodes" split by table
  bool order_dependent;               // a border set received both strands of one k-mer
  bool unable_to_backtrace;
  int n_path_lengths, path_lengths[2], reached_fuz;
  int forward_count;                  // count after the found check (before the all-paths recount)
  int rand_calls;
};

class Oracle {
 public:
  explicit Oracle(const Graph &g) : g_(g) {}
  // src/Gap2Seq.cpp:858-1556.  Uses libc rand() exactly like the reference.
  FillResult fill_gap(const std::string &kmer_left, const std::string &kmer_right, int gap_len, int k, int gap_err,
                      int left_max_fuz, int right_max_fuz, bool skip_confident, bool all_paths) const;

 private:
  const Graph &g_;
};

struct RunOptions {
  int k = 31, d_err = 500, max_fuz = 10;
  bool skip_confident = false, unique_paths = false, all_paths = true;
  bool stats = false;  // also print the -DSTATS "Nodes: ... Nonzeros: ..." line
};

// One scaffold record through the scanner/splicer of execute() (Gap2Seq.cpp:329-432).
// Appends the stdout lines the reference would print to `log`.
std::string fill_record(const Oracle &o, const std::string &seq, const std::string &comment, const RunOptions &opt,
                        std::string &log, int &gapcount, int &filledgapcount, std::vector<FillResult> *details);

}  // namespace g2s_oracle
#endif
