// g2s_gatb_glue.h - INTEGRATION.md form 2: the reference keeps GATB's graph (construction, .h5 IO, Bloom filter, MPHF)
// and only Gap2Seq::fill_gap (src/Gap2Seq.hpp:74-76) is replaced.  One declaration, same parameter list as the member
// function, so that the reference-side change is the two call sites (src/Gap2Seq.cpp:252, 380) and one #include.
#pragma once
#include <string>

#include <gatb/gatb_core.hpp>

struct subgraph_stats;

int g2s_fill_gap_graph(const Graph &graph, const std::string &kmer_left, const std::string &kmer_right, int gap_len, int k,
                       int gap_err, int left_max_fuz, int right_max_fuz, int *left_fuz, int *right_fuz, long long max_mem,
                       char *fill, bool skip_confident, bool all_paths, struct subgraph_stats *substats);
