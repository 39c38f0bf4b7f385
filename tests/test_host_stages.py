"""Host-only pieces of the product that the kernels hand their results to (gap2seq_b200/csrc/g2s_pool.h, g2s_finish.h),
exercised through the emulation library: thread pool, the window of libc's rand() stream computed ahead (RandAhead)."""
import ctypes

import emul_lib


def test_rand_window_reproduces_libc_and_commit_positions_it():
    """RandAhead::enabled() runs the self test (values computed ahead == rand(), commit(n) leaves libc after n calls) for
    window sizes below, at and far above the generator's 31-word table."""
    L = emul_lib.lib()
    assert L.emul_rand_skip_enabled() == 1
    assert L.emul_rand_ahead_enabled() == 1


def test_application_rand_stream_is_untouched_by_the_self_tests():
    libc = ctypes.CDLL(None)
    L = emul_lib.lib()
    libc.srand(99)
    want = [libc.rand() for _ in range(5)]
    libc.srand(99)
    L.emul_rand_ahead_enabled()
    L.emul_rand_skip_enabled()
    assert [libc.rand() for _ in range(5)] == want


def test_sequential_and_threaded_host_stages_agree(case_inputs, monkeypatch):
    """G2S_EMUL_THREADS=1 selects the sequential host stages; the default pool must give the same bytes and leave libc's
    generator in the same state (cfg2 has uncertain tracebacks, cfg3 general subgraphs)."""
    import subprocess, sys, os, json
    code = (
        "import sys, json, ctypes; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')\n"
        "import cases, emul_lib, oracle_lib\n"
        "out = {}\n"
        "for name in ['cfg2_small', 'cfg3_small', 'cfg1_multigap']:\n"
        "    case = cases.CASE_BY_NAME[name]; w = cases.build(case)\n"
        "    g = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)\n"
        "    eng = emul_lib.EmulEngine(w.k, g.kmers())\n"
        "    oracle_lib.srand(3)\n"
        "    r = eng.fill_records([x.seq for x in w.gaps], [x.name for x in w.gaps], dist_error=w.dist_error, fuz=w.fuz)\n"
        "    out[name] = [r[0], r[1], r[2], r[3], [ctypes.CDLL(None).rand() for _ in range(3)]]\n"
        "print(json.dumps(out))\n")
    res = []
    for threads in ("1", "4"):
        env = dict(os.environ, G2S_EMUL_THREADS=threads)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        res.append(json.loads(p.stdout.strip().splitlines()[-1]))
    assert res[0] == res[1]


def test_record_scanner_edge_cases_match_the_oracle_driver(case_inputs):
    """The gap driver (g2s_scaffold.h: wave scan, per-record replay, splice, statistics lines) against the oracle's
    restatement of Gap2Seq::execute() on records built to hit the scanner's corners: several gaps per record, gaps closer
    than k + 2 fuz (dependent flanks), lower-case n, N at the very start, left flanks shorter than k, gaps of length 1,
    records without gaps, an empty record."""
    import numpy as np
    import oracle_lib
    case, w, g = case_inputs("cfg1_small")
    eng = emul_lib.EmulEngine(w.k, g.kmers())
    rng = np.random.default_rng(5)
    genome = w.genome.tobytes().decode()
    recs = []
    for r in range(40):
        a = int(rng.integers(0, len(genome) - 4000))
        s = list(genome[a:a + int(rng.integers(200, 2500))])
        for _ in range(int(rng.integers(0, 4))):
            L = int(rng.choice([1, 2, 30, 45, 120, 300]))
            if len(s) - L - w.k <= 0:
                continue
            # a right flank shorter than k is undefined behaviour in the reference (negative right_max_fuz) and fenced
            # in the engine (DESIGN.md section 2 iv): every gap here keeps >= k characters to its right
            p = int(rng.integers(0, len(s) - L - w.k))
            ch = "n" if rng.random() < 0.3 else "N"
            s[p:p + L] = ch * L
        recs.append("".join(s))
    recs += ["", genome[100:400], "N" * 50 + genome[500:900], genome[2000:2020] + "N" * 60 + genome[2080:2500],
             genome[3000:3300] + "N" * 100 + genome[3400:3460] + "n" * 80 + genome[3540:3900]]
    names = [f"rec{i} x" for i in range(len(recs))]
    for kw in (dict(), dict(all_paths=False), dict(unique=True), dict(skip_confident=True)):
        oracle_lib.srand(2)
        filled, log, gc, fc, _ = eng.fill_records(recs, names, dist_error=300, fuz=10, **kw)
        oracle_lib.srand(2)
        ofilled, olog, ogc, ofc = [], "", 0, 0
        for s, n in zip(recs, names):
            f, l, a, b = g.fill_record(s, n, d_err=300, fuz=10, **kw)
            ofilled.append(f); olog += l; ogc += a; ofc += b
        assert (gc, fc) == (ogc, ofc)
        assert filled == ofilled
        assert log == olog


def test_emulated_device_traceback_equals_the_host_walk():
    """The kernel's traceback (g2s_fill.cuh: single-chain trellis -> fill string with case, no pack records) runs in the host
    emulation as well; with G2S_NO_DEVICE_TRACE=1 every gap is packed and walked on the host.  Same bytes either way."""
    import json, os, subprocess, sys
    code = (
        "import sys, json; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')\n"
        "import cases, emul_lib, oracle_lib\n"
        "from gap2seq_b200.scaffold import record_gaps\n"
        "out = {}\n"
        "for name in ['cfg1_small', 'cfg5_small', 'cfg2_small', 'cfg1_allupper']:\n"
        "    case = cases.CASE_BY_NAME[name]; w = cases.build(case); o = case.opts()\n"
        "    g = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)\n"
        "    eng = emul_lib.EmulEngine(w.k, g.kmers())\n"
        "    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, 500, skip_confident=o['all_upper'])]\n"
        "    oracle_lib.srand(4)\n"
        "    res, fills = eng.fill_batch(specs)\n"
        "    out[name] = [[(r.count, r.left_fuz, r.right_fuz, r.rand_calls, r.fill_len, list(r.substats.astuple())) for r in res], fills,\n"
        "                 sum(1 for r in res if r.status & 32)]\n"
        "print(json.dumps(out))\n")
    res = []
    for host_only in (False, True):
        env = dict(os.environ)
        env.pop("G2S_NO_DEVICE_TRACE", None)
        if host_only:
            env["G2S_NO_DEVICE_TRACE"] = "1"
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        res.append(json.loads(p.stdout.strip().splitlines()[-1]))
    for name in res[0]:
        assert res[0][name][:2] == res[1][name][:2], name
        assert res[1][name][2] == 0
    assert res[0]["cfg1_small"][2] > 0 and res[0]["cfg5_small"][2] > 0 and res[0]["cfg1_allupper"][2] > 0
