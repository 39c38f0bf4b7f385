// Gap2Seq-core - command line front end of the B200 engine with the reference tool's flags
// (reference src/Gap2Seq.cpp:51-91 + the GATB Tool options -nb-cores -verbose -version -help,
// gatb-core/src/gatb/tools/misc/impl/Tool.cpp:48-65; single-dash long names).  The Python wrapper of the
// reference (src/Gap2Seq.py:224-243) can drive this binary unchanged:
//
//   Gap2Seq-core -k 31 -solid 2 -reads R1.fa[,R2.fa] -scaffolds gaps.fa -filled out.fa
//                [-dist-error 500] [-fuz 10] [-max-mem 20] [-all-upper] [-best-only] [-unique] [-randseed S]
//                [-nb-cores N] [-device D] [-gpus G]
//   Gap2Seq-core ... -left L -right R -length n -filled out.fa            (single-gap mode, src/Gap2Seq.cpp:227-283)
//   Gap2Seq-core ... -gap-list gaps.tsv -filled out.fa                    (many single-gap runs at once, see below)
//
// Differences, all deliberate: the de Bruijn graph is built on the GPU from the reads (exact solid k-mer set, no
// Bloom filter, no ./.h5 file is written - an existing <reads>.h5 of GATB is loaded like the reference does); -nb-cores only sets the host threads of the subgraph analysis; -max-mem bounds the
// device scratch; -gpus G shards the gaps over GPUs D .. D+G-1 of the box (one k-mer table per GPU, replicated once;
// results gathered in input order, one rand() stream - the output does not depend on G).  stdout carries the same parameter echo, per-gap "Scaffold:" / "SubgraphStats:" lines and the final
// "Filled X gaps out of Y" as the reference.
#include <malloc.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/g2s.h"
#include "g2s_scaffold.h"

static bool read_sequences(const std::string &path, std::vector<std::string> &names, std::vector<std::string> &seqs) {
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

static void usage() {
  std::cout << "Gap2Seq-core (B200 engine)\n"
               "   -length     (1 arg) :    Length of a single gap  [default '']\n"
               "   -right      (1 arg) :    Right flank of a single gap  [default '']\n"
               "   -left       (1 arg) :    Left flank of a single gap  [default '']\n"
               "   -randseed   (1 arg) :    Random seed (0 to use current time)  [default '0']\n"
               "   -unique     (0 arg) :    If specified, only gaps with a unique path of best length are filled.\n"
               "   -best-only  (0 arg) :    If specified, only paths that have optimal length are considered.\n"
               "   -all-upper  (0 arg) :    If specified, all filled bases are in upper case.\n"
               "   -max-mem    (1 arg) :    Maximum memory usage of DP table computation in gigabytes (excluding DBG)  [default '20']\n"
               "   -fuz        (1 arg) :    Number of nucleotides to ignore on gap fringes  [default '10']\n"
               "   -dist-error (1 arg) :    Maximum error in gap estimates  [default '500']\n"
               "   -filled     (1 arg) :    FASTA file of filled scaffolds\n"
               "   -scaffolds  (1 arg) :    FASTA/Q file of scaffolds  [default '']\n"
               "   -reads      (1 arg) :    FASTA/Q files of reads. For several files use a comma separated list.\n"
               "   -solid      (1 arg) :    Threshold for solid k-mers  [default '2']\n"
               "   -k          (1 arg) :    k-mer length  [default '31']\n"
               "   -nb-cores   (1 arg) :    number of cores  [default '0']\n"
               "   -device     (1 arg) :    CUDA device ordinal  [default '0']\n"
               "   -gpus       (1 arg) :    number of GPUs the gaps are sharded over (devices D .. D+G-1)  [default '1']\n"
               "   -verbose    (1 arg) :    verbosity level  [default '1']\n"
               "   -version    (0 arg) :    version\n"
               "   -help       (0 arg) :    help\n";
}

int main(int argc, char **argv) {
  // this process hands megabytes of strings around per call: keep freed memory instead of returning it to the kernel
  mallopt(M_TRIM_THRESHOLD, 1 << 30); mallopt(M_MMAP_THRESHOLD, 1 << 28);
  std::map<std::string, std::string> opt = {{"-k", "31"}, {"-solid", "2"}, {"-dist-error", "500"}, {"-fuz", "10"},
                                            {"-max-mem", "20"}, {"-randseed", "0"}, {"-nb-cores", "0"}, {"-device", "0"}, {"-gpus", "1"},
                                            {"-verbose", "1"}, {"-scaffolds", ""}};
  std::map<std::string, bool> flag = {{"-all-upper", false}, {"-best-only", false}, {"-unique", false}};
  std::map<std::string, bool> saw;
  const char *one[] = {"-k", "-solid", "-reads", "-scaffolds", "-filled", "-dist-error", "-fuz", "-max-mem", "-randseed",
                       "-left", "-right", "-length", "-nb-cores", "-device", "-gpus", "-verbose", "-gap-list"};
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "-help" || a == "-h") { usage(); return 0; }
    if (a == "-version") { std::cout << "Gap2Seq-core B200 engine, ABI " << g2s_abi_version() << std::endl; return 0; }
    if (flag.count(a)) { flag[a] = true; continue; }
    bool ok = false;
    for (const char *o : one) if (a == o) ok = true;
    if (!ok || i + 1 >= argc) { std::cout << "ERROR: Unknown parameter '" << a << "'" << std::endl; usage(); return 1; }
    opt[a] = argv[++i];
    saw[a] = true;
  }
  if ((!saw["-reads"] && !saw["-gap-list"]) || !saw["-filled"]) { std::cout << "ERROR: Option '-reads' and '-filled' are mandatory" << std::endl; usage(); return 1; }

  const int k = atoi(opt["-k"].c_str()), solid = atoi(opt["-solid"].c_str());
  const int d_err = atoi(opt["-dist-error"].c_str()), max_fuz = atoi(opt["-fuz"].c_str());
  const long long max_mem = (long long)(atof(opt["-max-mem"].c_str()) * 1024 * 1024 * 1024);
  const bool skip_confident = flag["-all-upper"], unique_paths = flag["-unique"], all_paths = !flag["-best-only"];
  const int randseed = atoi(opt["-randseed"].c_str());
  const std::string reads = opt["-reads"], filled_path = opt["-filled"];
  srand(randseed > 0 ? (unsigned)randseed : (unsigned)time(NULL));   // src/Gap2Seq.cpp:178

  std::cout << "k-mer size: " << k << std::endl;
  std::cout << "Solidity threshold: " << solid << std::endl;
  std::cout << "Reads file: " << reads << std::endl;
  std::cout << "Filled scaffolds file: " << filled_path << std::endl;
  std::cout << "Distance error: " << d_err << std::endl;
  std::cout << "Max Fuz: " << max_fuz << std::endl;
  std::cout << "Max memory: " << max_mem << std::endl;
  std::cout << "Skip confident: " << skip_confident << std::endl;
  std::cout << "Unique: " << unique_paths << std::endl;
  std::cout << "All paths: " << all_paths << std::endl;
  std::cout << "Random seed: " << randseed << std::endl;

  g2s_config cfg{};
  cfg.abi_version = G2S_ABI_VERSION;
  cfg.device = atoi(opt["-device"].c_str());
  cfg.k = k;
  cfg.max_fuz = max_fuz;
  cfg.max_mem_bytes = max_mem;
  cfg.host_threads = atoi(opt["-nb-cores"].c_str());
  g2s_engine *eng = nullptr;
  const int n_gpus = std::max(1, atoi(opt["-gpus"].c_str()));
  std::vector<int32_t> devices(n_gpus);
  for (int d = 0; d < n_gpus; d++) devices[d] = cfg.device + d;
  if (g2s_create_multi(&cfg, devices.data(), n_gpus, &eng) != G2S_OK) { std::cout << "EXCEPTION: " << g2s_last_error() << std::endl; return 1; }

  // ---- per-gap filtered-reads mode in one process.  The reference's wrapper starts one Gap2Seq-core
  // per gap - "-reads <the reads near the gap> -left L -right R -length n" - and GATB builds one graph per process
  // (src/Gap2Seq.py:132-220).  A line of the gap list is such an invocation: name, left flank, right flank, length,
  // reads file(s), tab separated.  All graphs are built in one device pass and all gaps filled in one pass, each against
  // its own graph; the traceback of every gap starts from a freshly seeded rand() like its own process would, so that
  // log lines and sequences equal those of the per-gap runs.
  if (saw["-gap-list"]) {
    if (n_gpus > 1) { std::cout << "ERROR: -gap-list runs on one GPU" << std::endl; return 1; }
    std::ifstream lf(opt["-gap-list"].c_str());
    if (!lf) { std::cout << "EXCEPTION: cannot read " << opt["-gap-list"] << std::endl; return 1; }
    struct Item { std::string name, left, right; int length; };
    std::vector<Item> items;
    std::vector<uint64_t> off(1, 0), first(1, 0);
    std::string bases, line;
    while (std::getline(lf, line)) {
      if (line.empty() || line[0] == '#') continue;
      std::vector<std::string> f;
      std::stringstream ls(line);
      std::string tok;
      while (std::getline(ls, tok, '\t')) f.push_back(tok);
      if (f.size() < 5) { std::cout << "EXCEPTION: gap list line needs name, left, right, length, reads: " << line << std::endl; return 1; }
      items.push_back(Item{f[0], f[1], f[2], atoi(f[3].c_str())});
      std::vector<std::string> rn, rs;
      std::stringstream fs(f[4]);
      std::string file;
      while (std::getline(fs, file, ',')) {
        if (!read_sequences(file, rn, rs)) { std::cout << "DBG building failed: cannot read " << file << std::endl; return 1; }
      }
      for (auto &r : rs) { bases += r; off.push_back(bases.size()); }
      first.push_back(off.size() - 1);
    }
    auto tb0 = std::chrono::steady_clock::now();
    if (g2s_build_graphs_from_reads(eng, bases.data(), off.data(), off.size() - 1, first.data(), items.size(), solid) != G2S_OK) {
      std::cout << "DBG building failed: " << g2s_last_error() << std::endl;
      return 1;
    }
    auto tb1 = std::chrono::steady_clock::now();
    std::vector<g2s_gap> gaps(items.size());
    std::vector<int> lmfs(items.size()), rmfs(items.size());
    std::vector<char> skip(items.size(), 0);
    size_t cap = 16;
    for (size_t i = 0; i < items.size(); i++) {
      const Item &it = items[i];
      g2s_gap g{};
      if ((int)it.left.length() < k || (int)it.right.length() < k) { skip[i] = 1; }   // "Flanks need to be at least k length"
      const int lmf = skip[i] ? 0 : std::min((int)it.left.length() - k, max_fuz), rmf = skip[i] ? 0 : std::min((int)it.right.length() - k, max_fuz);
      lmfs[i] = lmf; rmfs[i] = rmf;
      // like g2s_fill_gap: fill_gap takes the left k-mers and the targets from the BEGINNING of its flank strings and
      // starts the right pass from the END of the right one (src/Gap2Seq.cpp:871-885, 1000, 1119)
      g.kmer_left = it.left.c_str();
      g.kmer_right = it.right.c_str();
      g.kmer_right_starts = (!skip[i] && (int)it.right.length() > k + rmf) ? it.right.c_str() + (it.right.length() - k - rmf) : nullptr;
      g.gap_len = it.length; g.gap_err = d_err; g.left_max_fuz = lmf; g.right_max_fuz = rmf;
      g.flags = (skip_confident ? G2S_GAP_SKIP_CONFIDENT : 0) | (all_paths ? G2S_GAP_ALL_PATHS : 0);
      g.graph = (int32_t)i;
      gaps[i] = g;
      cap = std::max(cap, (size_t)it.length + k + d_err + lmf + rmf + 3);
    }
    std::vector<g2s_gap> run;
    std::vector<size_t> run_of;
    for (size_t i = 0; i < items.size(); i++) if (!skip[i]) { run_of.push_back(i); run.push_back(gaps[i]); }
    g2s_batch *batch = nullptr;
    if (g2s_submit(eng, run.data(), run.size(), &batch) != G2S_OK) { std::cout << "EXCEPTION: " << g2s_last_error() << std::endl; return 1; }
    auto tb2 = std::chrono::steady_clock::now();
    std::ofstream output(filled_path.c_str());
    if (!output) { std::cout << "EXCEPTION: cannot write " << filled_path << std::endl; return 1; }
    std::vector<char> buf(cap, 0);
    size_t at = 0;
    int filledcount = 0;
    for (size_t i = 0; i < items.size(); i++) {
      const Item &it = items[i];
      if (skip[i]) { std::cerr << "Flanks need to be at least k length" << std::endl; continue; }
      srand(randseed > 0 ? (unsigned)randseed : (unsigned)time(NULL));   // every gap is its own run (src/Gap2Seq.cpp:178)
      g2s_result r{};
      buf[0] = 0;
      if (g2s_batch_trace(batch, at++, &r, buf.data()) != G2S_OK) { std::cout << "EXCEPTION: " << g2s_last_error() << std::endl; return 1; }
      const int lmf = lmfs[i], rmf = rmfs[i];
      std::string log;
      g2s::format_statistics(log, (int)it.left.length() - lmf - r.left_fuz, (int)it.left.length(), (int)it.left.length() + it.length, r.count,
                             buf.data(), (int)strlen(buf.data()), k, lmf, rmf, r.left_fuz, r.right_fuz, skip_confident, unique_paths,
                             r.substats, it.length, "");
      std::cout << log;
      std::string filledSeq;
      if (r.count > 0 && (!unique_paths || r.count == 1)) { filledSeq = it.left.substr(0, it.left.length() - r.left_fuz) + buf.data(); filledcount++; }
      else filledSeq = it.left + std::string((size_t)it.length, 'N') + it.right;
      output << ">" << it.name << "\n" << filledSeq << "\n";
    }
    g2s_batch_free(batch);
    std::cout << "Filled " << filledcount << " gaps out of " << items.size() << std::endl;
    std::cerr << "g2s: " << items.size() << " graphs built in " << std::chrono::duration<double>(tb1 - tb0).count() << " s, gaps filled in "
              << std::chrono::duration<double>(tb2 - tb1).count() << " s" << std::endl;
    g2s_destroy(eng);
    return 0;
  }

  // ---- de Bruijn graph from the reads (stand-in for Graph::create, src/Gap2Seq.cpp:199-218); timed separately
  auto t0 = std::chrono::steady_clock::now();
  const std::string reads_graph = reads + ".h5";
  if (std::ifstream(reads_graph.c_str()).good()) {
    // a GATB graph file next to the reads is loaded instead (src/Gap2Seq.cpp:171, 195-197): its dsk/solid partitions are
    // the node set; they are read from the HDF5 container directly and imported into the device table
    std::cout << "Loading from " << reads_graph << std::endl;
    if (g2s_import_h5(eng, reads_graph.c_str()) != G2S_OK) { std::cout << "EXCEPTION: " << g2s_last_error() << std::endl; return 1; }
  } else {
    // the reads files go to the device as they are and are parsed there (g2s_build_from_read_files); text that is neither
    // FASTA nor FASTQ is reported by the engine
    std::vector<std::string> files;
    std::stringstream ss(reads);
    std::string file;
    while (std::getline(ss, file, ',')) files.push_back(file);
    std::vector<const char *> fp;
    for (auto &f : files) fp.push_back(f.c_str());
    if (g2s_build_from_read_files(eng, fp.data(), fp.size(), solid, nullptr) != G2S_OK) {
      std::cout << "DBG building failed: " << g2s_last_error() << std::endl;
      return 1;
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  std::cerr << "g2s: " << g2s_num_kmers(eng) << " solid k-mers on the device, graph "
            << std::chrono::duration<double>(t1 - t0).count() << " s" << std::endl;

  std::ofstream output(filled_path.c_str());
  if (!output) { std::cout << "EXCEPTION: cannot write " << filled_path << std::endl; return 1; }

  if (saw["-left"] && saw["-right"] && saw["-length"]) {   // single-gap mode, src/Gap2Seq.cpp:227-283
    const std::string left = opt["-left"], right = opt["-right"];
    const int length = atoi(opt["-length"].c_str());
    if ((int)left.length() < k || (int)right.length() < k) { std::cerr << "Flanks need to be at least k length" << std::endl; return 0; }
    const int lmf = std::min((int)left.length() - k, max_fuz), rmf = std::min((int)right.length() - k, max_fuz);
    std::vector<char> buf((size_t)length + k + d_err + lmf + rmf + 3, 0);
    int lf = 0, rf = 0;
    g2s_subgraph_stats ss{};
    int n = g2s_fill_gap(eng, left.c_str(), right.c_str(), length, k, d_err, lmf, rmf, &lf, &rf, max_mem, buf.data(),
                         skip_confident, all_paths, &ss);
    if (n < -1) { std::cout << "EXCEPTION: " << g2s_last_error() << std::endl; return 1; }
    const char *fill = &buf[lmf - lf];
    std::string log;
    g2s::format_statistics(log, (int)left.length() - lmf - lf, (int)left.length(), (int)left.length() + length, n, fill,
                           (int)strlen(fill), k, lmf, rmf, lf, rf, skip_confident, unique_paths, ss, length, "");
    std::cout << log;
    std::string filledSeq;
    if (n > 0 && (!unique_paths || n == 1)) filledSeq = left.substr(0, left.length() - lf) + fill;
    else filledSeq = left + std::string((size_t)length, 'N') + right;
    output << ">" << "\n" << filledSeq << "\n";
    g2s_destroy(eng);
    return 0;
  }

  std::vector<std::string> names, seqs;
  if (!read_sequences(opt["-scaffolds"], names, seqs)) { std::cout << "EXCEPTION: cannot read " << opt["-scaffolds"] << std::endl; return 1; }
  std::vector<const char *> sp(seqs.size()), cp(seqs.size());
  for (size_t i = 0; i < seqs.size(); i++) { sp[i] = seqs[i].c_str(); cp[i] = names[i].c_str(); }
  std::vector<char *> filled(seqs.size() + 1, nullptr);
  char *log = nullptr;
  int32_t gapcount = 0, filledcount = 0;
  g2s_run_options ro{k, d_err, max_fuz, skip_confident, unique_paths, all_paths};
  auto t2 = std::chrono::steady_clock::now();
  if (g2s_fill_records(eng, sp.data(), cp.data(), seqs.size(), &ro, filled.data(), &log, &gapcount, &filledcount, nullptr,
                       nullptr) != G2S_OK) {
    std::cout << "EXCEPTION: " << g2s_last_error() << std::endl;
    return 1;
  }
  auto t3 = std::chrono::steady_clock::now();
  std::cout << log;
  for (size_t i = 0; i < seqs.size(); i++) { output << ">" << names[i] << "\n" << filled[i] << "\n"; g2s_free(filled[i]); }
  g2s_free(log);
  output.flush();
  std::cout << "Filled " << filledcount << " gaps out of " << gapcount << std::endl;   // src/Gap2Seq.cpp:437
  g2s_timing tm{};
  g2s_get_timing(eng, &tm);
  std::cerr << "g2s: fill " << std::chrono::duration<double>(t3 - t2).count() << " s (kernels " << tm.kernel_ms
            << " ms, host " << tm.host_ms << " ms, " << tm.launches << " launches)" << std::endl;
  g2s_destroy(eng);
  return 0;
}
