// g2s_gapcutter.cpp - GapCutter of this package: turns scaffolds into a contigs file, a gaps file (flank + N run +
// flank records for Gap2Seq-core) and a BED file, so that gaps can be filled independently and merged back by
// GapMerger.  The file formats and the record names (" scaffold S contig C gap G [split 1|2 k]") are the reference
// tool's (src/GapCutter.cpp:47-50 markers, :193-311 output), so that either implementation can feed the other's
// merger; tests/test_cut_merge.py compares all three files byte for byte with the reference binary.
// Host-only.
//
// Structure (not the reference's scanning loop): a scaffold is first tokenised ONCE into its alternating runs of
// bases ("islands") and Ns ("lakes").  A planner then walks the island list with a cursor and classifies what starts
// at the cursor into one of six outcomes (Plan::Kind); each outcome says how many characters it consumes and which
// slices of the scaffold become a gap record, a contig record and BED lines.  The emitter only copies slices.
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

struct Island {          // a maximal run of bases and the lake (run of N / n) that follows it; either may be empty
  size_t at, bases, lake;
  size_t lake_at() const { return at + bases; }
  size_t end() const { return at + bases + lake; }
};

std::vector<Island> tokenise(const std::string &s) {
  std::vector<Island> v;
  size_t i = 0;
  const size_t n = s.size();
  auto is_n = [&](size_t p) { return s[p] == 'N' || s[p] == 'n'; };
  while (i < n || v.empty()) {
    Island is{i, 0, 0};
    while (i < n && !is_n(i)) i++;
    is.bases = i - is.at;
    while (i < n && is_n(i)) i++;
    is.lake = i - is.at - is.bases;
    v.push_back(is);
    if (i >= n) break;
  }
  return v;
}

struct Slice { size_t at, len; };

struct Plan {
  enum Kind {
    REST_IS_CONTIG,      // no fillable gap ahead: everything from the cursor to the end of the scaffold is one contig
    LAKE_STAYS,          // the lake keeps its place inside a contig (no room for a flank on one of its sides)
    ONE_GAP,             // flank + lake + flank
    TWO_GAPS_ONE_ISLAND, // two lakes share the island between them: "split 1" / "split 2 k" records
    MASKED_GAP           // several lakes around islands too short to be flanks, masked into one gap of n
  } kind;
  size_t consumed;       // characters of the scaffold this outcome accounts for, counted from the cursor
  Slice contig;          // contig record
  Slice gap_left, gap_right;   // ONE_GAP / MASKED_GAP: record = left flank .. (lake | n x masked) .. right flank
  size_t masked;         // MASKED_GAP: characters replaced by n between the flanks
  Slice second;          // TWO_GAPS_ONE_ISLAND: the second record (k bases of the shared island + lake + flank)
};

struct Rules { size_t k, fuz; bool mask, split; };

// What starts at `cursor` (a position inside island `i`, or at its lake when the island is used up)?
Plan plan_at(const std::vector<Island> &isl, size_t i, size_t cursor, size_t total, const Rules &R) {
  const Island &cur = isl[i];
  const size_t head = cur.lake_at() - cursor;              // bases between the cursor and the next lake
  const size_t flank_cap = R.k + R.fuz;
  Plan p{};
  p.kind = Plan::REST_IS_CONTIG; p.consumed = total - cursor; p.contig = Slice{cursor, total - cursor};
  if (cur.lake == 0) return p;                             // last island
  if (head < R.k) {                                        // no left flank
    p.kind = Plan::LAKE_STAYS; p.consumed = head + cur.lake; p.contig = Slice{cursor, p.consumed};
    return p;
  }
  const size_t lf = std::min(head, flank_cap);
  const Slice left{cur.lake_at() - lf, lf};
  const Slice kept{cursor, head - lf};                     // what the contig keeps of the head
  const Island *nx = i + 1 < isl.size() ? &isl[i + 1] : nullptr;
  const size_t mid = nx ? nx->bases : 0;                   // island behind the lake
  const bool mid_is_last = !nx || nx->lake == 0;
  const size_t beyond = (!mid_is_last && i + 2 < isl.size()) ? isl[i + 2].bases : 0;   // island behind the next lake
  auto one_gap = [&](size_t rf) {
    p.kind = Plan::ONE_GAP; p.contig = kept; p.gap_left = Slice{left.at, lf + cur.lake}; p.gap_right = Slice{cur.end(), rf};
    p.consumed = head + cur.lake + rf;
  };
  // the island behind the lake can serve as right flank and still leave a left flank for what follows,
  // or it is the end of the scaffold
  if (mid >= 2 * R.k) { one_gap(std::min(mid, flank_cap)); return p; }
  if (mid >= R.k && (mid_is_last || beyond == 0)) { one_gap(mid); return p; }
  if (mid >= R.k) {
    // a k..2k-1 island between two lakes: either both gaps share it, or only the first gap is cut here
    if (R.split && beyond >= R.k) {
      const size_t rf2 = std::min(beyond, flank_cap);
      p.kind = Plan::TWO_GAPS_ONE_ISLAND; p.contig = kept;
      p.gap_left = Slice{left.at, lf + cur.lake}; p.gap_right = Slice{nx->at, mid};
      p.second = Slice{nx->lake_at() - R.k, R.k + nx->lake + rf2};
      p.consumed = head + cur.lake + mid + nx->lake + rf2;
    } else {
      one_gap(std::min(mid, flank_cap));
    }
    return p;
  }
  // the island behind the lake is too short for a flank: look for the first island of at least k bases further on
  size_t j = i + 1;
  while (j < isl.size() && isl[j].bases < R.k && isl[j].lake > 0) j++;
  // isl[j] is the candidate right flank; it does not qualify if the scaffold ended in short islands / a lake
  const bool have_right = j < isl.size() && j > i + 1 && isl[j].bases >= R.k;
  if (!have_right) {
    // the reference walks lake by lake and stops at a non-empty island shorter than k only when nothing follows
    if (j < isl.size() && isl[j].bases >= R.k) { /* unreachable: mid < k means j > i + 1 */ }
    return p;                                              // REST_IS_CONTIG
  }
  const size_t span = isl[j].at - cur.lake_at();           // lakes and short islands between the flanks
  if (R.mask) {
    const size_t rf = std::min(isl[j].bases, flank_cap);
    p.kind = Plan::MASKED_GAP; p.contig = kept; p.gap_left = Slice{left.at, lf}; p.gap_right = Slice{isl[j].at, rf};
    p.masked = span; p.consumed = head + span + rf;
  } else {
    p.kind = Plan::LAKE_STAYS; p.consumed = head + span; p.contig = Slice{cursor, p.consumed};
  }
  return p;
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
  const Rules R{(size_t)atoi(opt["-k"].c_str()), (size_t)atoi(opt["-fuz"].c_str()), mask, !no_split};

  std::cout << "Scaffolds file: " << opt["-scaffolds"] << std::endl;
  std::cout << "Contigs file: " << opt["-contigs"] << std::endl;
  std::cout << "Gaps file: " << opt["-gaps"] << std::endl;
  std::cout << "BED file: " << opt["-bed"] << std::endl;
  std::cout << "k-mer size: " << R.k << std::endl;
  std::cout << "Fuz: " << R.fuz << std::endl;
  std::cout << "Mask: " << R.mask << std::endl;
  std::cout << "Split: " << R.split << std::endl;

  std::vector<std::string> names, seqs;
  if (!g2s::read_sequences(opt["-scaffolds"], names, seqs)) { std::cout << "EXCEPTION: cannot read " << opt["-scaffolds"] << std::endl; return EXIT_FAILURE; }
  g2s::FastaWriter contigs(opt["-contigs"]), gaps(opt["-gaps"]);
  std::ofstream bed(opt["-bed"].c_str());

  int n_contig = 0, n_gap = 0, n_scaffold = 0;
  for (size_t r = 0; r < seqs.size(); r++) {
    const std::string &seq = seqs[r];
    const size_t sp = names[r].find(' ');
    if (sp == std::string::npos) { std::cout << "EXCEPTION: scaffold header without a space: " << names[r] << std::endl; return EXIT_FAILURE; }
    const std::string chrom = names[r].substr(0, sp);
    const std::vector<Island> isl = tokenise(seq);
    auto text = [&](const Slice &s) { return seq.substr(s.at, s.len); };
    auto bed_line = [&](size_t from, size_t to) { bed << chrom << "\t" << from << "\t" << to << std::endl; };
    size_t cursor = 0, i = 0;
    while (cursor < seq.size()) {
      while (isl[i].end() <= cursor && i + 1 < isl.size()) i++;   // island that holds the cursor (or whose lake starts at it)
      const Plan p = plan_at(isl, i, cursor, seq.size(), R);
      const std::string tag = names[r] + " scaffold " + std::to_string(n_scaffold) + " contig " + std::to_string(n_contig);
      const std::string gap_tag = tag + " gap " + std::to_string(n_gap);
      switch (p.kind) {
        case Plan::REST_IS_CONTIG:
        case Plan::LAKE_STAYS:
          contigs.insert(tag, text(p.contig));
          break;
        case Plan::ONE_GAP:
          gaps.insert(gap_tag, text(p.gap_left) + text(p.gap_right));
          contigs.insert(gap_tag, text(p.contig));
          bed_line(p.gap_left.at, p.gap_right.at + p.gap_right.len);
          n_gap++;
          break;
        case Plan::TWO_GAPS_ONE_ISLAND:
          gaps.insert(gap_tag + " split 1", text(p.gap_left) + text(p.gap_right));
          gaps.insert(gap_tag + " split 2 " + std::to_string(R.k), text(p.second));
          contigs.insert(gap_tag, text(p.contig));
          bed_line(p.gap_left.at, p.gap_right.at + p.gap_right.len);
          bed_line(p.gap_right.at, p.second.at + p.second.len);
          n_gap++;
          break;
        case Plan::MASKED_GAP:
          gaps.insert(gap_tag, text(p.gap_left) + std::string(p.masked, 'n') + text(p.gap_right));
          contigs.insert(gap_tag, text(p.contig));
          bed_line(p.gap_left.at, p.gap_right.at + p.gap_right.len);
          n_gap++;
          break;
      }
      n_contig++;
      cursor += p.consumed;
    }
    gaps.flush();
    contigs.flush();
    n_scaffold++;
  }
  std::cout << "Cut " << n_scaffold << " scaffolds into " << n_contig << " contigs and " << n_gap << " gaps" << std::endl;
  return EXIT_SUCCESS;
}
