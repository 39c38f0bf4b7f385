// C entry points over g2s_oracle for ctypes (tests/, smoke(), bench.py cpu_baseline).
// TEST INFRASTRUCTURE ONLY - see g2s_oracle.hpp.
#include <cstdlib>
#include <cstring>

#include "g2s_oracle.hpp"

using namespace g2s_oracle;

extern "C" {

struct orc_result {
  int32_t count, left_fuz, right_fuz, forward_count;
  int32_t n_path_lengths, path_len0, path_len1, reached_fuz;
  int32_t order_dependent, unable_to_backtrace, substats_valid, rand_calls;
  uint64_t substats[6];
  uint64_t cells_right, cells_left, nodes_right, nodes_left;
  int32_t fill_len, pad;
};

void *orc_graph_from_reads(int k, int solid, const char *bases, const uint64_t *offsets, size_t n_reads) {
  std::vector<std::string> reads(n_reads);
  for (size_t i = 0; i < n_reads; i++) reads[i].assign(bases + offsets[i], bases + offsets[i + 1]);
  Graph *g = new Graph();
  g->build_from_reads(k, solid, reads);
  return g;
}

void *orc_graph_from_kmers(int k, const uint64_t *kmers, size_t n) {
  Graph *g = new Graph();
  g->build_from_kmers(k, kmers, n);
  return g;
}

size_t orc_graph_size(void *g) { return ((Graph *)g)->size(); }

void orc_graph_kmers(void *g, uint64_t *out, uint8_t *adj) {
  std::vector<uint64_t> v;
  ((Graph *)g)->sorted_kmers(v);
  for (size_t i = 0; i < v.size(); i++) {
    out[i] = v[i];
    if (adj) adj[i] = ((Graph *)g)->adjacency(v[i]);
  }
}

void orc_graph_free(void *g) { delete (Graph *)g; }

void orc_srand(unsigned seed) { srand(seed); }

static void pack(const FillResult &r, orc_result *out, char *fill, size_t fill_cap) {
  memset(out, 0, sizeof(*out));
  out->count = r.count; out->left_fuz = r.left_fuz; out->right_fuz = r.right_fuz;
  out->forward_count = r.forward_count;
  out->n_path_lengths = r.n_path_lengths; out->path_len0 = r.path_lengths[0]; out->path_len1 = r.path_lengths[1];
  out->reached_fuz = r.reached_fuz;
  out->order_dependent = r.order_dependent; out->unable_to_backtrace = r.unable_to_backtrace;
  out->substats_valid = r.substats_valid; out->rand_calls = r.rand_calls;
  out->substats[0] = r.substats.vertices; out->substats[1] = r.substats.edges;
  out->substats[2] = r.substats.nontrivial_components; out->substats[3] = r.substats.size_nontrivial_components;
  out->substats[4] = r.substats.vertices_final; out->substats[5] = r.substats.edges_final;
  out->cells_right = r.cells_right; out->cells_left = r.cells_left;
  out->nodes_right = r.nodes_right; out->nodes_left = r.nodes_left;
  out->fill_len = (int32_t)r.fill.size();
  if (fill && fill_cap) {
    size_t n = r.fill.size() < fill_cap - 1 ? r.fill.size() : fill_cap - 1;
    memcpy(fill, r.fill.data(), n);
    fill[n] = 0;
  }
}

// One call of Gap2Seq::fill_gap (src/Gap2Seq.hpp:74-76); consumes libc rand() like the reference.
int orc_fill_gap(void *g, const char *left, const char *right, int gap_len, int k, int gap_err, int lmf, int rmf,
                 int skip_confident, int all_paths, orc_result *out, char *fill, size_t fill_cap) {
  Oracle o(*(Graph *)g);
  FillResult r = o.fill_gap(left, right, gap_len, k, gap_err, lmf, rmf, skip_confident != 0, all_paths != 0);
  pack(r, out, fill, fill_cap);
  return r.count;
}

// One scaffold record through the execute() scanner.  Returns malloc'd strings the caller frees with orc_free.
char *orc_fill_record(void *g, const char *seq, const char *comment, int k, int d_err, int max_fuz, int skip_confident,
                      int unique_paths, int all_paths, int stats, char **log_out, int *gapcount, int *filledcount) {
  Oracle o(*(Graph *)g);
  RunOptions opt;
  opt.k = k; opt.d_err = d_err; opt.max_fuz = max_fuz;
  opt.skip_confident = skip_confident != 0; opt.unique_paths = unique_paths != 0; opt.all_paths = all_paths != 0;
  opt.stats = stats != 0;
  std::string log;
  int gc = 0, fc = 0;
  std::string filled = fill_record(o, seq, comment, opt, log, gc, fc, nullptr);
  if (gapcount) *gapcount += gc;
  if (filledcount) *filledcount += fc;
  if (log_out) *log_out = strdup(log.c_str());
  return strdup(filled.c_str());
}

void orc_free(void *p) { free(p); }

}  // extern "C"
