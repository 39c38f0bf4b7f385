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
