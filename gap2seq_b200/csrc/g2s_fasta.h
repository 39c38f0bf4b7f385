// g2s_fasta.h - minimal FASTA/FASTQ reading and single-line FASTA writing for the command line tools of this
// package (Gap2Seq-core, GapCutter, GapMerger).  Output format = what GATB's BankFasta::insert writes
// (gatb-core/src/gatb/bank/impl/BankFasta.cpp:230-270): ">comment\n" and the whole sequence on one line; the file is
// created when the first record is written.
#pragma once
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

namespace g2s {

inline bool read_sequences(const std::string &path, std::vector<std::string> &names, std::vector<std::string> &seqs) {
  std::ifstream f(path.c_str());
  if (!f) return false;
  std::string line;
  bool fastq = false, first = true;
  int phase = 0;
  while (std::getline(f, line)) {
    if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
    if (first && !line.empty()) { fastq = line[0] == '@'; first = false; }
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

class FastaWriter {   // opened lazily, like BankFasta: no record, no file
 public:
  explicit FastaWriter(const std::string &path) : path_(path) {}
  ~FastaWriter() { if (f_) fclose(f_); }
  FastaWriter(const FastaWriter &) = delete;
  FastaWriter &operator=(const FastaWriter &) = delete;
  bool insert(const std::string &comment, const std::string &seq) {
    if (!f_) f_ = fopen(path_.c_str(), "w");
    if (!f_) return false;
    fprintf(f_, ">%s\n", comment.c_str());
    fwrite(seq.data(), 1, seq.size(), f_);
    fputc('\n', f_);
    return true;
  }
  void flush() { if (f_) fflush(f_); }

 private:
  std::string path_;
  FILE *f_ = nullptr;
};

}  // namespace g2s
