// Command-line front end of the oracle with Gap2Seq-core's flag names (src/Gap2Seq.cpp:51-91).
// TEST INFRASTRUCTURE ONLY - see g2s_oracle.hpp.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "g2s_oracle.hpp"

using namespace g2s_oracle;

static bool read_fasta(const std::string &path, std::vector<std::string> &names, std::vector<std::string> &seqs) {
  std::ifstream f(path.c_str());
  if (!f) return false;
  std::string line;
  bool fastq = false;
  int phase = 0;
  while (std::getline(f, line)) {
    if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
    if (names.empty() && seqs.empty() && !line.empty() && line[0] == '@') fastq = true;
    if (fastq) {
      if (phase == 0) { names.push_back(line.substr(1)); seqs.push_back(""); }
      else if (phase == 1) seqs.back() = line;
      phase = (phase + 1) & 3;
      continue;
    }
    if (!line.empty() && line[0] == '>') { names.push_back(line.substr(1)); seqs.push_back(""); }
    else if (!seqs.empty()) seqs.back() += line;
  }
  return true;
}

int main(int argc, char **argv) {
  RunOptions opt;
  int solid = 2, randseed = 0;
  std::string reads, scaffolds, filled;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> const char * { return (i + 1 < argc) ? argv[++i] : ""; };
    if (a == "-k") opt.k = atoi(next());
    else if (a == "-solid") solid = atoi(next());
    else if (a == "-reads") reads = next();
    else if (a == "-scaffolds") scaffolds = next();
    else if (a == "-filled") filled = next();
    else if (a == "-dist-error") opt.d_err = atoi(next());
    else if (a == "-fuz") opt.max_fuz = atoi(next());
    else if (a == "-randseed") randseed = atoi(next());
    else if (a == "-all-upper") opt.skip_confident = true;
    else if (a == "-best-only") opt.all_paths = false;
    else if (a == "-unique") opt.unique_paths = true;
    else if (a == "-stats") opt.stats = true;
    else if (a == "-nb-cores" || a == "-max-mem" || a == "-verbose") next();
    else { fprintf(stderr, "g2s_oracle: unknown option %s\n", a.c_str()); return 2; }
  }
  srand(randseed > 0 ? (unsigned)randseed : (unsigned)time(NULL));   // Gap2Seq.cpp:178

  std::vector<std::string> rn, rs;
  std::string one;
  size_t start = 0;
  while (start <= reads.size()) {   // comma separated list, Gap2Seq.cpp:199-210
    size_t c = reads.find(',', start);
    one = reads.substr(start, c == std::string::npos ? std::string::npos : c - start);
    if (!one.empty() && !read_fasta(one, rn, rs)) { fprintf(stderr, "cannot read %s\n", one.c_str()); return 1; }
    if (c == std::string::npos) break;
    start = c + 1;
  }
  auto t0 = std::chrono::steady_clock::now();
  Graph g;
  g.build_from_reads(opt.k, solid, rs);
  auto t1 = std::chrono::steady_clock::now();
  std::vector<std::string>().swap(rs);
  Oracle o(g);

  std::vector<std::string> sn, ss;
  if (!read_fasta(scaffolds, sn, ss)) { fprintf(stderr, "cannot read %s\n", scaffolds.c_str()); return 1; }
  std::ofstream out(filled.c_str());
  int gapcount = 0, filledcount = 0;
  for (size_t i = 0; i < ss.size(); i++) {
    std::string log;
    std::string f = fill_record(o, ss[i], sn[i], opt, log, gapcount, filledcount, nullptr);
    std::cout << log;
    out << ">" << sn[i] << "\n" << f << "\n";
  }
  auto t2 = std::chrono::steady_clock::now();
  std::cout << "Filled " << filledcount << " gaps out of " << gapcount << std::endl;   // Gap2Seq.cpp:437
  fprintf(stderr, "g2s_oracle: solid k-mers %zu, graph %.3f s, fill %.3f s\n", g.size(),
          std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
  return 0;
}
