// g2s_gapmerger.cpp - GapMerger of this package: puts the contigs and the (filled) gap records that GapCutter wrote
// back together into scaffolds.  A contig whose name carries " gap G" is followed by gap record G; a gap that was
// split in two (" split 1" / " split 2 <k>") contributes its first part whole and its second part without the k
// characters both parts share.  Record names lose everything from " scaffold " on.  Same behaviour, flags and output
// as the reference tool (src/GapMerger.cpp:153-232; tests/test_cut_merge.py compares byte for byte), but the gap
// records are indexed once instead of being rescanned for every contig (the reference is quadratic in the number
// of gaps).  Host-only.
//
//   GapMerger -scaffolds <out fa> -contigs <fa> -gaps <filled fa>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "g2s_fasta.h"

namespace {

// the integer that follows `marker` in a record name and ends at `until` (or at the end of the name); -1 without marker
int index_after(const std::string &name, const std::string &marker, const std::string &until) {
  const size_t at = name.find(marker);
  if (at == std::string::npos) return -1;
  const size_t from = at + marker.size();
  const size_t to = until.empty() ? std::string::npos : name.find(until);
  return std::stoi(to == std::string::npos ? name.substr(from) : name.substr(from, to - from));
}

struct GapParts {          // what a scan of the gaps file for one index collects before it stops
  std::string first, second;
  bool done = false;
};

}  // namespace

int main(int argc, char **argv) {
  std::map<std::string, std::string> opt;
  for (int ai = 1; ai < argc; ai++) {
    const std::string a = argv[ai];
    if ((a == "-scaffolds" || a == "-contigs" || a == "-gaps" || a == "-nb-cores" || a == "-verbose") && ai + 1 < argc) opt[a] = argv[++ai];
    else { std::cout << "Unknown parameter '" << a << "'" << std::endl; return EXIT_FAILURE; }
  }
  for (const char *need : {"-scaffolds", "-contigs", "-gaps"})
    if (!opt.count(need)) { std::cout << "Option '" << need << "' is mandatory" << std::endl; return EXIT_FAILURE; }
  std::cout << "Scaffolds file: " << opt["-scaffolds"] << std::endl;
  std::cout << "Contigs file: " << opt["-contigs"] << std::endl;
  std::cout << "Gaps file: " << opt["-gaps"] << std::endl;

  std::vector<std::string> cnames, cseqs, gnames, gseqs;
  if (!g2s::read_sequences(opt["-contigs"], cnames, cseqs)) { std::cout << "EXCEPTION: cannot read " << opt["-contigs"] << std::endl; return EXIT_FAILURE; }
  g2s::read_sequences(opt["-gaps"], gnames, gseqs);   // no gaps file: nothing was cut

  // gap index -> what a scan of the file from its beginning collects for that index (an unsplit record, or the two
  // parts of a split gap)
  std::unordered_map<int, GapParts> by_index;
  for (size_t g = 0; g < gnames.size(); g++) {
    const int idx = index_after(gnames[g], " gap ", " split ");
    if (idx < 0) continue;
    GapParts &gp = by_index[idx];
    if (gp.done) continue;                                 // the scan stopped before this record
    const size_t sp = gnames[g].find(" split ");
    if (sp == std::string::npos) { gp.first = gseqs[g]; gp.done = true; continue; }   // an unsplit record ends the scan
    const int part = std::stoi(gnames[g].substr(sp + 7, 1));
    if (part == 1) gp.first = gseqs[g];
    else {
      const size_t shared = (size_t)std::stoi(gnames[g].substr(sp + 7 + 2));   // characters the two parts have in common
      gp.second = shared <= gseqs[g].size() ? gseqs[g].substr(shared) : std::string();
    }
    if (!gp.first.empty() && !gp.second.empty()) gp.done = true;
  }

  g2s::FastaWriter out(opt["-scaffolds"]);
  auto emit = [&](const std::string &name, const std::string &seq) {
    const size_t at = name.find(" scaffold ");
    out.insert(at == std::string::npos ? name : name.substr(0, at), seq);
  };
  int n_contigs = 0, n_gaps = 0, cur = 0;
  std::string scaffold, scaffold_name = cnames.empty() ? std::string() : cnames[0];
  for (size_t c = 0; c < cnames.size(); c++) {
    const int s_idx = index_after(cnames[c], " scaffold ", " contig ");
    const int g_idx = index_after(cnames[c], " gap ", " split ");
    n_contigs++;
    if (s_idx != cur) {   // the previous scaffold is complete
      emit(scaffold_name, scaffold);
      scaffold.clear();
      scaffold_name = cnames[c];
      cur = s_idx;
    }
    scaffold += cseqs[c];
    if (g_idx != -1) {
      auto it = by_index.find(g_idx);
      if (it != by_index.end()) scaffold += it->second.first + it->second.second;
      n_gaps++;
    }
  }
  if (!scaffold.empty()) { emit(scaffold_name, scaffold); cur++; }
  out.flush();
  std::cout << "Merged " << n_contigs << " contigs and " << n_gaps << " gaps into " << cur << " scaffolds" << std::endl;
  return EXIT_SUCCESS;
}
