"""Workload.reads_near(): the bounded sample bench.py hands to the CPU reference arm for cfg4 (a prefix of the gap list
with only the reads that overlap those gaps' DP windows) fills to exactly the same records and log lines as the full
read set - so the sample measures the same per-gap work as the full workload."""
import oracle_lib
from gap2seq_b200 import synth


def test_reads_near_gaps_give_the_same_fills_as_all_reads():
    w = synth.make_workload("cfg4", 0.02)
    sub = w.gaps[:40]
    near = w.reads_near(sub)
    assert 0 < near.shape[0] < w.reads.shape[0] // 3
    full = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
    part = oracle_lib.OracleGraph.from_reads(w.k, w.solid, near)
    oracle_lib.srand(1)
    a = [full.fill_record(g.seq, g.name, d_err=w.dist_error, fuz=w.fuz) for g in sub]
    oracle_lib.srand(1)
    b = [part.fill_record(g.seq, g.name, d_err=w.dist_error, fuz=w.fuz) for g in sub]
    assert a == b
    assert all(f[0][100:-100] == g.truth for f, g in zip(a, sub))   # error-free reads: the fill is the removed sequence
