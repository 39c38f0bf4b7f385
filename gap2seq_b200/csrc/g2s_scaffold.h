// g2s_scaffold.h - the gap driver around the engine: the scaffold scanner / splicer of Gap2Seq::execute()
// (reference src/Gap2Seq.cpp:329-432) and print_statistics (src/Gap2Seq.cpp:100-156), restructured so that
// all gaps of all records go to the device in a few large batches instead of one fill_gap call at a time.
//
// The reference scans a record left to right and a gap's flanks can depend on the PREVIOUS gap's result
// (prevGapEnd moves by right_fuz when that gap was filled).  Phase 1 therefore submits gaps in waves: a gap is
// submitted as soon as its parameters are decided, which is immediately unless the previous attempted gap of the
// same record ends less than k + 2*fuz bases before it.  Phase 2 then replays the exact reference control flow in
// record order, tracing each attempted gap (that is where libc rand() is consumed, so the rand() stream is the
// one of `Gap2Seq-core -nb-cores 1`), printing the statistics lines and splicing - including the reference's
// prefix-dropping `filledSeq = ...` assignment for multi-gap records (src/Gap2Seq.cpp:396).
#pragma once
#include <atomic>
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "g2s_finish.h"

namespace g2s {

// a caller's NUL-terminated string, not copied (records are megabytes per batch)
struct StrRef {
  const char *p = "";
  size_t n = 0;
  StrRef() {}
  StrRef(const char *s) : p(s ? s : ""), n(s ? strlen(s) : 0) {}
  StrRef(const std::string &s) : p(s.data()), n(s.size()) {}
  size_t size() const { return n; }
  const char *data() const { return p; }
  char operator[](size_t i) const { return p[i]; }
  std::string substr(size_t pos, size_t len = std::string::npos) const {
    if (pos > n) pos = n;
    return std::string(p + pos, std::min(len, n - pos));
  }
  std::string str() const { return std::string(p, n); }
};

struct RunOptions {
  int k = 31, d_err = 500, max_fuz = 10;
  bool skip_confident = false, unique_paths = false, all_paths = true;
};

// device pass for a list of gaps -> detached host batch (no rand() consumed)
typedef std::function<int(const std::vector<g2s_gap> &, HostBatch &)> SubmitFn;

struct GapReq {
  int kmer_start, lmf, rmf, gap;
  size_t i_after;          // index just after the N run
  int batch = -1, idx = -1;
};

struct RecScan {
  size_t i = 0;
  long long prev = 0;          // prevGapEnd when known
  int last_attempt = -1;       // request whose result decides prev, -1 when prev is exact
  std::vector<GapReq> reqs;
  bool done = false;
};

inline bool is_gap_char(char c) { return c == 'N' || c == 'n'; }
// index of the next N / n at or after i (seq.size() if none).  The records are NUL-terminated strings, so libc's
// vectorised strcspn does the scan; an embedded NUL only ends the fast scan early and the byte loop takes over.
inline size_t next_gap_char(const StrRef &seq, size_t i) {
  const size_t n = seq.size();
  while (i < n) {
    i += strcspn(seq.data() + i, "Nn");
    if (i >= n || is_gap_char(seq[i])) break;
    i++;   // a NUL inside the record
  }
  return i < n ? i : n;
}

// src/Gap2Seq.cpp:100-156
inline void format_statistics(std::string &log, int filledStart, int gapStart, int gapEnd, int paths, const char *fill,
                              int fill_len, int k, int lmf, int rmf, int lf, int rf, bool skip_confident, bool unique_paths,
                              const g2s_subgraph_stats &ss, int gap, const StrRef &comment_ref) {
  const std::string comment = comment_ref.str();
  char buf[1024];
  if (paths > 0 && (!unique_paths || paths == 1)) {
    const int filledLen = fill_len - k;
    if (!skip_confident) {
      // fills hold letters only (ACGT / acgt): upper case = bit 5 clear; a plain sum the compiler vectorises
      int upper = 0;
      for (int j = 0; j < filledLen; j++) upper += ((unsigned char)fill[j] & 0x20u) == 0;
      const int lower = filledLen - upper;
      snprintf(buf, sizeof buf,
               " GapStart: %d GapEnd: %d GapLength: %d PathsFound: %d FilledStart: %d FilledEnd: %d FilledGapLength: %d"
               " LeftMaxFuz: %d LeftFuz: %d RightMaxFuz: %d RightFuz: %d ConfidentBases: %d TotalBases: %d\n",
               gapStart, gapEnd, gap, paths, filledStart, filledStart + filledLen, filledLen, lmf, lf, rmf, rf, upper,
               upper + lower);
      log += "Scaffold: " + comment + buf;
      snprintf(buf, sizeof buf,
               "SubgraphStats: Vertices: %llu Edges: %llu NontrivialStrongComponents: %llu"
               " SizeNontrivialStrongComponents: %llu VerticesFinal: %llu EdgesFinal: %llu\n",
               (unsigned long long)ss.vertices, (unsigned long long)ss.edges, (unsigned long long)ss.nontrivial_components,
               (unsigned long long)ss.size_nontrivial_components, (unsigned long long)ss.vertices_final,
               (unsigned long long)ss.edges_final);
      log += buf;
    } else {
      snprintf(buf, sizeof buf,
               " GapStart: %d GapEnd: %d GapLength: %d PathsFound: %d FilledStart: %d FilledEnd: %d FilledGapLength: %d"
               " LeftFuz: %d RightFuz: %d\n",
               gapStart, gapEnd, gap, paths, filledStart, filledStart + filledLen, filledLen, lf, rf);
      log += "Scaffold: " + comment + buf;
    }
  } else {
    snprintf(buf, sizeof buf,
             " GapStart: %d GapEnd: %d GapLength: %d PathsFound: 0 FilledStart: 0 FilledEnd: 0 FilledGapLength: 0"
             " LeftMaxFuz: %d LeftFuz: %d RightMaxFuz: %d RightFuz: %d%s\n",
             gapStart, gapEnd, gap, lmf, lf, rmf, rf, paths == -1 ? " Memory limit exceeded" : "");
    log += "Scaffold: " + comment + buf;
  }
}

struct ScaffoldRun {
  const std::vector<StrRef> &seqs;
  const std::vector<StrRef> &comments;
  RunOptions opt;
  std::vector<RecScan> recs;
  std::vector<std::unique_ptr<HostBatch>> batches;
  double phase1_ms = 0, phase2_ms = 0, trace_ms = 0, wait_ms = 0, flush_ms = 0;  // wall time of the two phases of run() (profiling aid)
  ThreadPool *pool = nullptr;                         // host threads for the tracebacks and the per-record replay
  RandAhead own_ahead, *aheadp = &own_ahead;          // the libc rand() values the tracebacks will consume; callers that
  RawBuf<char> own_arena, *arenap = &own_arena;       // run many times hand in persistent ones (no page faults per call)
  char **filled_c = nullptr;                          // when set: record r's filled sequence goes to a malloc'ed C string
                                                      // filled_c[r] (one copy) and `filled[r]` stays empty

  ScaffoldRun(const std::vector<StrRef> &s, const std::vector<StrRef> &c, const RunOptions &o)
      : seqs(s), comments(c), opt(o), recs(s.size()) {}
  ~ScaffoldRun() { if (pool) pool->wait_async(); }

  bool filled_status(int count) const { return count > 0 && (!opt.unique_paths || count == 1); }

  // Phase 1 helper: scan record r as far as its gaps are decided, appending requests to `out`.
  void scan(size_t r, std::vector<std::pair<size_t, int>> &out) {
    const StrRef &seq = seqs[r];
    RecScan &rs = recs[r];
    const int k = opt.k, fuz = opt.max_fuz;
    while (rs.i < seq.size()) {
      if (!is_gap_char(seq[rs.i])) { rs.i = next_gap_char(seq, rs.i + 1); continue; }
      const size_t gs = rs.i;
      bool independent = false;
      if (rs.last_attempt >= 0) {
        const GapReq &pr = rs.reqs[rs.last_attempt];
        if (pr.batch >= 0 && batches[pr.batch]) {
          const HostBatch &hb = *batches[pr.batch];
          const int cnt = hb.peek_count(pr.idx);
          rs.prev = (long long)pr.i_after + (filled_status(cnt) ? hb.peek_right_fuz(pr.idx) : 0);
          rs.last_attempt = -1;
        } else if ((long long)gs - (long long)pr.i_after >= (long long)k + 2 * fuz) {
          independent = true;   // left flank cannot reach back to wherever the previous gap ends
        } else {
          return;               // suspended until the previous gap's batch has run
        }
      }
      GapReq q;
      q.lmf = independent ? fuz : (int)std::min<long long>((long long)gs + k - rs.prev, fuz);
      q.kmer_start = (int)gs - k - q.lmf;
      size_t i = gs;
      int gap = 0;
      while (i < seq.size() && is_gap_char(seq[i])) { i++; gap++; }
      q.gap = gap;
      q.i_after = i;
      q.rmf = (int)std::min<long long>((long long)seq.size() - ((long long)i + k), fuz);
      // a right flank shorter than k is undefined behaviour in the reference (negative right_max_fuz): skipped here
      bool ok = q.rmf >= 0 && i + k + q.rmf <= seq.size();
      for (int j = 0; j < k + q.rmf && ok; j++) if (is_gap_char(seq[i + j])) ok = false;
      const bool left_ok = independent || (long long)q.kmer_start >= rs.prev;
      if (left_ok && ok) {
        rs.reqs.push_back(q);
        rs.last_attempt = (int)rs.reqs.size() - 1;
        out.push_back(std::make_pair(r, rs.last_attempt));
      } else if (!independent) {
        rs.prev = (long long)i;
        rs.last_attempt = -1;
      } else {
        // independent but right flank invalid: the gap is skipped whatever the previous result is
        rs.prev = (long long)i;
        rs.last_attempt = -1;
      }
      rs.i = i;
    }
    rs.done = true;
  }

  int run(const SubmitFn &submit, std::vector<std::string> &filled, std::string &log, int &gapcount, int &filledcount,
          std::vector<g2s_result> *details) {
    const int k = opt.k;
    const auto t_begin = std::chrono::steady_clock::now();
    size_t rand_bound = 0;
    if (pool) aheadp->begin();
    // ---- phase 1: device waves
    for (;;) {
      std::vector<std::pair<size_t, int>> wave;
      bool any_open = false;
      if (pool && batches.empty() && recs.size() >= 16) {   // first wave: nothing to look up, records scan in parallel
        std::vector<std::vector<std::pair<size_t, int>>> part(recs.size());
        pool->parallel_for(recs.size(), 4, [&](size_t r) { scan(r, part[r]); });
        for (size_t r = 0; r < recs.size(); r++) {
          wave.insert(wave.end(), part[r].begin(), part[r].end());
          if (!recs[r].done) any_open = true;
        }
      } else {
        for (size_t r = 0; r < recs.size(); r++) {
          if (recs[r].done) continue;
          scan(r, wave);
          if (!recs[r].done) any_open = true;
        }
      }
      if (wave.empty()) { if (any_open) { if (pool) pool->wait_async(); return -1; } break; }
      const int b = (int)batches.size();
      batches.emplace_back(new HostBatch());
      std::vector<g2s_gap> gaps(wave.size());
      for (size_t j = 0; j < wave.size(); j++) {
        GapReq &q = recs[wave[j].first].reqs[wave[j].second];
        const StrRef &seq = seqs[wave[j].first];
        q.batch = b; q.idx = (int)j;
        g2s_gap &g = gaps[j];
        g.kmer_left = seq.data() + q.kmer_start;   // k + lmf / k + rmf characters inside the record (not NUL terminated)
        g.kmer_right = seq.data() + q.i_after;
        g.kmer_right_starts = nullptr;
        g.gap_len = q.gap; g.gap_err = opt.d_err; g.left_max_fuz = q.lmf; g.right_max_fuz = q.rmf;
        g.flags = (opt.skip_confident ? G2S_GAP_SKIP_CONFIDENT : 0) | (opt.all_paths ? G2S_GAP_ALL_PATHS : 0);
        g.graph = 0;
      }
      // the rand() values this wave's tracebacks can consume are computed while the device pass runs
      // (a filling path is about as long as the gap; the rare longer ones extend the window after the pass, see trace_many)
      for (const g2s_gap &g : gaps) rand_bound += (size_t)g.gap_len + g.left_max_fuz + g.right_max_fuz + 64;
      if (pool && aheadp->ok()) {
        pool->wait_async();
        const size_t want = rand_bound + 1;
        pool->run_async([this, want] { aheadp->extend(want); });
      }
      int rc = submit(gaps, *batches.back());
      if (rc != 0) { if (pool) pool->wait_async(); return rc; }
      if (!any_open) {
        bool all_done = true;
        for (auto &rs : recs) if (!rs.done) all_done = false;
        if (all_done) break;
      }
    }
    // ---- phase 2: the reference's control flow, record by record (src/Gap2Seq.cpp:329-432).  The tracebacks of all
    // attempted gaps run first, on all host threads, at their positions in the libc rand() stream (trace_many); the
    // replay below then only formats and splices, which is independent per record.
    const auto t_mid = std::chrono::steady_clock::now();
    phase1_ms = std::chrono::duration<double, std::milli>(t_mid - t_begin).count();
    if (pool) pool->wait_async();
    wait_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_mid).count();
    std::vector<TraceJob> jobs;
    std::vector<size_t> first_job(seqs.size() + 1, 0);
    for (size_t r = 0; r < seqs.size(); r++) {
      first_job[r] = jobs.size();
      for (const GapReq &q : recs[r].reqs) jobs.push_back(TraceJob{batches[q.batch].get(), (size_t)q.idx});
    }
    first_job[seqs.size()] = jobs.size();
    std::vector<g2s_result> res;
    std::vector<size_t> fill_off;
    std::vector<const char *> fill_at;   // device-traced fills are read where the DMA left them, not copied first
    trace_many(jobs, res, *arenap, fill_off, pool, aheadp->ok() ? aheadp : nullptr, &fill_at);
    trace_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_mid).count();

    filled.assign(seqs.size(), std::string());
    std::vector<std::string> rlog(seqs.size());
    std::vector<int> rgc(seqs.size(), 0), rfc(seqs.size(), 0);
    std::atomic<int> bad{0};
    auto one_record = [&](size_t r) {
      const StrRef &seq = seqs[r];
      const StrRef &comment = comments[r];
      static thread_local std::string tl_seq;   // reused from record to record when the output is a C string
      std::string own_seq;
      std::string &filledSeq = filled_c ? tl_seq : own_seq;
      filledSeq.clear();
      filledSeq.reserve(seq.size() + 64);
      std::string &lg = rlog[r];
      long long prev = 0;
      size_t i = 0, next_req = 0;
      while (i < seq.size()) {
        if (!is_gap_char(seq[i])) { i = next_gap_char(seq, i + 1); continue; }
        rgc[r]++;
        const int lmf = (int)std::min<long long>((long long)i + k - prev, opt.max_fuz);
        const int kmer_start = (int)i - k - lmf;
        int gap = 0;
        while (i < seq.size() && is_gap_char(seq[i])) { i++; gap++; }
        const int rmf = (int)std::min<long long>((long long)seq.size() - ((long long)i + k), opt.max_fuz);
        bool ok = rmf >= 0 && i + k + rmf <= seq.size();
        for (int j = 0; j < k + rmf && ok; j++) if (is_gap_char(seq[i + j])) ok = false;
        if ((long long)kmer_start >= prev && ok) {
          if (next_req >= recs[r].reqs.size()) { bad = -2; return; }
          const GapReq &q = recs[r].reqs[next_req];
          if (q.kmer_start != kmer_start || q.lmf != lmf || q.rmf != rmf || q.gap != gap || q.i_after != i) { bad = -3; return; }
          const size_t jb = first_job[r] + next_req++;
          const g2s_result &rs = res[jb];
          const char *fill = fill_at[jb];
          const int s = rs.count;
          const int filledStart = (int)filledSeq.length() + kmer_start + k + lmf - rs.left_fuz - (int)prev;
          const int gapStart = kmer_start + k + lmf;
          format_statistics(lg, filledStart, gapStart, (int)i, s, fill, rs.fill_len, k, lmf, rmf, rs.left_fuz,
                            rs.right_fuz, opt.skip_confident, opt.unique_paths, rs.substats, gap, comment);
          if (filled_status(s)) {
            rfc[r]++;
            // '=' not '+=': the prefix of a multi-gap record is dropped, as in the reference
            filledSeq.assign(seq.data() + prev, (size_t)(kmer_start + k + lmf - rs.left_fuz - prev));
            filledSeq.append(fill, (size_t)rs.fill_len);
            filledSeq.resize(filledSeq.size() - k);   // drop the right k-mer
            i += rs.right_fuz;
          } else {
            filledSeq.append(seq.data() + prev, (size_t)(kmer_start + k + lmf + gap - prev));
          }
        } else {
          filledSeq.append(seq.data() + prev, (size_t)(kmer_start + k + lmf + gap - prev));
        }
        prev = (long long)i;
      }
      if (next_req != recs[r].reqs.size()) { bad = -4; return; }
      filledSeq.append(seq.data() + prev, seq.size() - (size_t)prev);
      if (filled_c) {
        char *o = (char *)malloc(filledSeq.size() + 1);
        if (o) memcpy(o, filledSeq.c_str(), filledSeq.size() + 1);
        filled_c[r] = o;
      } else {
        filled[r].swap(filledSeq);
      }
    };
    if (pool) pool->parallel_for(seqs.size(), 8, one_record);
    else for (size_t r = 0; r < seqs.size(); r++) one_record(r);
    if (bad.load() != 0) return bad.load();
    const auto t_fl = std::chrono::steady_clock::now();
    for (auto &b : batches) if (b) b->flush_staging(pool);   // after the replay: it read the fills out of the staging
    flush_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_fl).count();
    size_t lbytes = 0;
    for (const std::string &l : rlog) lbytes += l.size();
    log.reserve(log.size() + lbytes);
    for (size_t r = 0; r < seqs.size(); r++) { log += rlog[r]; gapcount += rgc[r]; filledcount += rfc[r]; }
    if (details) details->insert(details->end(), res.begin(), res.end());
    phase2_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_mid).count();
    return 0;
  }
};

}  // namespace g2s
