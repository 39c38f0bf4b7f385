// g2s_gapcutter.cpp - GapCutter of this package: cuts scaffolds at their N runs into a contigs file and a gaps file
// (flank + gap + flank records for Gap2Seq-core) plus a BED file, so that gaps can be filled independently and merged
// back by GapMerger.  Behaviour, flags, record names (" scaffold S contig C gap G [split 1|2 k]") and file formats
// follow the reference tool (src/GapCutter.cpp:193-311) so that either implementation can feed the other's merger;
// tests/test_cut_merge.py compares the outputs byte for byte with the reference binary.  Host-only.
//
//   GapCutter -scaffolds <fa> -contigs <out fa> -gaps <out fa> -bed <out bed> [-k 31] [-fuz 10] [-mask] [-no-split]
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "g2s_fasta.h"

namespace {

bool is_n(char c) { return c == 'N' || c == 'n'; }
// length of the run of non-N (want_n = false) or N (want_n = true) characters starting at `from`
size_t run_length(const std::string &s, size_t from, bool want_n) {
  if (from >= s.size()) return 0;
  size_t pos = from;
  while (pos < s.size() && is_n(s[pos]) == want_n) pos++;
  return pos - from;
}

}  // namespace

int main(int argc, char **argv) {
  std::map<std::string, std::string> opt = {{"-k", "31"}, {"-fuz", "10"}};
  bool mask = false, no_split = false;
  for (int ai = 1; ai < argc; ai++) {
    const std::string a = argv[ai];
    if (a == "-mask") mask = true;
    else if (a == "-no-split") no_split = true;
    else if ((a == "-scaffolds" || a == "-contigs" || a == "-gaps" || a == "-bed" || a == "-k" || a == "-fuz" || a == "-nb-cores" ||
              a == "-verbose") && ai + 1 < argc) opt[a] = argv[++ai];
    else { std::cout << "Unknown parameter '" << a << "'" << std::endl; return EXIT_FAILURE; }
  }
  for (const char *need : {"-scaffolds", "-contigs", "-gaps", "-bed"})
    if (!opt.count(need)) { std::cout << "Option '" << need << "' is mandatory" << std::endl; return EXIT_FAILURE; }
  const size_t k = (size_t)atoi(opt["-k"].c_str()), fuz = (size_t)atoi(opt["-fuz"].c_str());

  std::cout << "Scaffolds file: " << opt["-scaffolds"] << std::endl;
  std::cout << "Contigs file: " << opt["-contigs"] << std::endl;
  std::cout << "Gaps file: " << opt["-gaps"] << std::endl;
  std::cout << "BED file: " << opt["-bed"] << std::endl;
  std::cout << "k-mer size: " << k << std::endl;
  std::cout << "Fuz: " << fuz << std::endl;
  std::cout << "Mask: " << mask << std::endl;
  std::cout << "Split: " << !no_split << std::endl;

  std::vector<std::string> names, seqs;
  if (!g2s::read_sequences(opt["-scaffolds"], names, seqs)) { std::cout << "EXCEPTION: cannot read " << opt["-scaffolds"] << std::endl; return EXIT_FAILURE; }
  g2s::FastaWriter contigs(opt["-contigs"]), gaps(opt["-gaps"]);
  std::ofstream bed(opt["-bed"].c_str());

  int n_contig = 0, n_gap = 0, n_scaffold = 0;
  for (size_t r = 0; r < seqs.size(); r++) {
    const std::string &seq = seqs[r];
    const size_t sp = names[r].find(' ');
    if (sp == std::string::npos) { std::cout << "EXCEPTION: scaffold header without a space: " << names[r] << std::endl; return EXIT_FAILURE; }
    const std::string scaffold_name = names[r].substr(0, sp);
    size_t pos = 0;
    while (pos < seq.size()) {
      const std::string comment = names[r] + " scaffold " + std::to_string(n_scaffold) + " contig " + std::to_string(n_contig);
      const std::string gap_comment = comment + " gap " + std::to_string(n_gap);
      const size_t before = run_length(seq, pos, false);          // sequence before the next gap
      const size_t gap1 = run_length(seq, pos + before, true);      // that gap
      if (before > 0 && gap1 == 0) {                              // no gap left: the rest is a contig
        contigs.insert(comment, seq.substr(pos));
        n_contig++;
        break;
      }
      if (before < k) {                                         // no room for a left flank: the gap stays in the contig
        contigs.insert(comment, seq.substr(pos, before + gap1));
        n_contig++;
        pos += before + gap1;
        continue;
      }
      const size_t lflank = std::min(before, k + fuz);
      const size_t middle = run_length(seq, pos + before + gap1, false);
      const size_t gap2 = run_length(seq, pos + before + gap1 + middle, true);
      const size_t after = run_length(seq, pos + before + gap1 + middle + gap2, false);
      const size_t rec_start = pos + before - lflank;                    // start of the gap record
      if (middle >= 2 * k || (middle >= k && after == 0)) {            // enough sequence on both sides
        const size_t rflank = middle >= 2 * k ? std::min(middle, k + fuz) : middle;   // a last gap takes the rest as right flank
        gaps.insert(gap_comment, seq.substr(rec_start, lflank + gap1 + rflank));
        contigs.insert(gap_comment, seq.substr(pos, before - lflank));
        bed << scaffold_name << "\t" << rec_start << "\t" << pos + before + gap1 + rflank << std::endl;
        n_gap++; n_contig++;
        pos += before + gap1 + rflank;
        continue;
      }
      if (middle >= k) {                                        // two gaps share the sequence between them
        if (!no_split && after >= k) {
          const size_t rflank2 = std::min(after, k + fuz);
          // the first gap gets the whole middle section as right flank, the second one only k characters of it
          gaps.insert(gap_comment + " split 1", seq.substr(rec_start, lflank + gap1 + middle));
          gaps.insert(gap_comment + " split 2 " + std::to_string(k), seq.substr(pos + before + gap1 + middle - k, k + gap2 + rflank2));
          contigs.insert(gap_comment, seq.substr(pos, before - lflank));
          bed << scaffold_name << "\t" << rec_start << "\t" << pos + before + gap1 + middle << std::endl;
          bed << scaffold_name << "\t" << pos + before + gap1 << "\t" << pos + before + gap1 + middle + gap2 + rflank2 << std::endl;
          n_gap++; n_contig++;
          pos += before + gap1 + middle + gap2 + rflank2;
        } else {
          const size_t rflank = std::min(middle, k + fuz);
          gaps.insert(gap_comment, seq.substr(rec_start, lflank + gap1 + rflank));
          contigs.insert(gap_comment, seq.substr(pos, before - lflank));
          bed << scaffold_name << "\t" << rec_start << "\t" << pos + before + gap1 + rflank << std::endl;
          n_gap++; n_contig++;
          pos += before + gap1 + rflank;
        }
        continue;
      }
      // gaps around sequence too short to be a flank: extend over all of them
      size_t span = gap1 + middle + gap2, tail = after;
      while (tail > 0 && tail < k) {
        span += tail + run_length(seq, pos + before + span + tail, true);
        tail = run_length(seq, pos + before + span, false);
      }
      if (tail < k) {                                         // nothing long enough follows: the rest is a contig
        contigs.insert(comment, seq.substr(pos));
        n_contig++;
        break;
      }
      if (mask) {                                           // everything between the flanks becomes one gap
        const size_t rflank = std::min(tail, k + fuz);
        gaps.insert(gap_comment, seq.substr(rec_start, lflank) + std::string(span, 'n') + seq.substr(pos + before + span, rflank));
        contigs.insert(gap_comment, seq.substr(pos, before - lflank));
        bed << scaffold_name << "\t" << rec_start << "\t" << pos + before + span + rflank << std::endl;
        n_gap++; n_contig++;
        pos += before + span + rflank;
      } else {
        contigs.insert(comment, seq.substr(pos, before + span));
        n_contig++;
        pos += before + span;
      }
    }
    gaps.flush();
    contigs.flush();
    n_scaffold++;
  }
  std::cout << "Cut " << n_scaffold << " scaffolds into " << n_contig << " contigs and " << n_gap << " gaps" << std::endl;
  return EXIT_SUCCESS;
}
