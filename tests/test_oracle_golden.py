"""The CPU oracle (oracle/g2s_oracle.cpp) against the golden outputs of the UNMODIFIED reference binary
(tests/golden/, made by tests/golden/make_golden.py with Gap2Seq-core -nb-cores 1 -randseed 1).
This is what pins the oracle: byte-identical filled FASTA, identical Scaffold:/SubgraphStats: lines and
identical -DSTATS Nodes/Nonzeros (DP cell) counts."""
import pytest

import cases
import oracle_lib


def run_oracle(case, w, g):
    o = case.opts()
    oracle_lib.srand(o["randseed"])
    out, log, gaps, filled = [], [], 0, 0
    for rec in w.gaps:
        f, l, gc, fc = g.fill_record(rec.seq, rec.name, d_err=o["dist_error"], fuz=o["fuz"],
                                     skip_confident=o["all_upper"], unique=o["unique"],
                                     all_paths=not o["best_only"], stats=True)
        out.append(f">{rec.name}\n{f}\n")
        log.append(l)
        gaps += gc
        filled += fc
    log.append(f"Filled {filled} gaps out of {gaps}\n")
    return "".join(out), "".join(log)


@pytest.mark.parametrize("name", [c.name for c in cases.CASES])
def test_oracle_matches_reference_golden(name, case_inputs):
    case, w, g = case_inputs(name)
    fa, log = run_oracle(case, w, g)
    gfa, glog = cases.golden_paths(name)
    assert fa == open(gfa).read()
    assert log == open(glog).read()


def test_known_answer_cell_counts(case_inputs):
    """SURVEY 3.2 anchor: in unique sequence Nonzeros = (rmf+ceil((g+e)/2)+1) + (g/2+e/2+lmf); for long gaps the
    forward table instead ends at rnode(rmf), i.e. at level lmf+g+k+rmf (the walk stays inside the right set)."""
    case, w, g = case_inputs("cfg1_small")
    for rec in w.gaps:
        i = rec.seq.index("N")
        gl = rec.gap_len
        left, right = rec.seq[i - 41:i], rec.seq[i + gl:i + gl + 41]
        res, fill = g.fill_gap(left, right, gl, 500, 10, 10)
        assert res.cells_right == 10 + -(-(gl + 500) // 2) + 1
        assert res.cells_left == max(gl // 2 + 250 + 10, 10 + gl + 31 + 10 + 1)
        assert res.count == 1 and fill[:-31] == rec.truth and fill.isupper()
        assert res.substats[0] == gl + 43 and res.substats[1] == gl + 42
