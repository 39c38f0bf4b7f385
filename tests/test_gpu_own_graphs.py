"""SURVEY 8f rank 2, per-gap filtered-reads mode: every gap has its own graph, built from the reads near it.  The
reference's wrapper runs one Gap2Seq-core process + one GATB construction per gap (src/Gap2Seq.py:132-220, single-gap
mode src/Gap2Seq.cpp:227-283); g2s_build_graphs_from_reads builds the graphs of a batch in one device pass and
g2s_fill_batch fills every gap against its own graph in one pass.  Checked against the oracle run gap by gap on a graph
of that gap's reads alone, same rand() stream."""
import numpy as np
import pytest

import oracle_lib
from gap2seq_b200 import engine
from gap2seq_b200.scaffold import record_gaps
from gap2seq_b200.synth import make_workload

pytestmark = pytest.mark.gpu


def _gap_specs(w, graphs=True):
    specs = []
    for gi, rec in enumerate(w.gaps):
        (g,) = record_gaps(rec.seq, k=w.k, fuz=w.fuz, dist_error=w.dist_error)
        g.graph = gi if graphs else 0
        specs.append(g)
    return specs


@pytest.mark.parametrize("name,scale,solid", [("cfg1", 1.0, 2), ("cfg2", 0.1, 2), ("cfg5", 0.01, 3)])
def test_each_gap_against_the_graph_of_its_own_reads(name, scale, solid):
    w = make_workload(name, scale)
    read_sets = [w.reads_near([rec], reach=150) for rec in w.gaps]
    specs = _gap_specs(w)
    graphs = [oracle_lib.OracleGraph.from_reads(w.k, solid, rs) for rs in read_sets]
    with engine.Engine(k=w.k) as e:
        e.build_graphs_from_reads(read_sets, solid)
        assert [e.graph_num_kmers(g) for g in range(len(read_sets))] == [G.size() for G in graphs]
        start, compared = 0, 0
        while start < len(specs):   # re-seed both sides behind an order-dependent gap (tests/parity.py)
            oracle_lib.srand(7 + start)
            res, fills = e.fill_batch(specs[start:])
            oracle_lib.srand(7 + start)
            nxt = len(specs)
            for j, (s, r, f) in enumerate(zip(specs[start:], res, fills)):
                o, of = graphs[s.graph].fill_gap(s.left, s.right, s.gap_len, s.gap_err, s.left_max_fuz, s.right_max_fuz)
                if o.order_dependent:
                    assert r.status & 1
                    nxt = start + j + 1
                    break
                got = (r.count, r.left_fuz, r.right_fuz, f, r.forward_count, r.rand_calls, r.cells_right, r.cells_left)
                want = (o.count, o.left_fuz, o.right_fuz, of, o.forward_count, o.rand_calls, o.cells_right, o.cells_left)
                assert got == want, (start + j, got[:3] + got[4:], want[:3] + want[4:])
                compared += 1
            start = nxt
    assert compared >= len(specs) - 2
    assert sum(1 for G in graphs if G.size() > 0) == len(graphs)


def test_graphs_do_not_leak_into_each_other():
    """Two gaps, reads swapped: each gap must see only its own (useless) graph and stay unfilled; with the right graphs
    both are filled."""
    w = make_workload("cfg1", 1.0)
    recs = w.gaps[:2]
    sets = [w.reads_near([r], reach=150) for r in recs]
    specs = _gap_specs(w)[:2]
    with engine.Engine(k=w.k) as e:
        e.build_graphs_from_reads(sets, 2)
        res, _ = e.fill_batch(specs)
        assert all(r.count > 0 for r in res)
        e.build_graphs_from_reads(sets[::-1], 2)
        res, _ = e.fill_batch(specs)
        assert all(r.count == 0 for r in res)
        with pytest.raises(engine.G2SError, match="graph out of range"):
            e.fill_batch([engine.GapSpec(specs[0].left, specs[0].right, specs[0].gap_len, graph=2)])
        # back to one graph for everything
        e.build_from_reads(w.reads, w.solid)
        res, _ = e.fill_batch(_gap_specs(w, graphs=False)[:2])
        assert all(r.count > 0 for r in res)


def test_empty_and_tiny_graphs():
    w = make_workload("cfg1", 1.0)
    specs = _gap_specs(w)[:3]
    sets = [w.reads_near([w.gaps[0]], reach=150), np.zeros((0, 100), dtype=np.uint8), w.reads[:1]]
    with engine.Engine(k=w.k) as e:
        e.build_graphs_from_reads(sets, 2)
        assert e.graph_num_kmers(1) == 0 and e.graph_num_kmers(2) == 0 and e.graph_num_kmers(0) > 0 and e.graph_num_kmers(3) == -1
        res, _ = e.fill_batch(specs)
        assert res[0].count > 0 and res[1].count == 0 and res[2].count == 0


def test_gap_list_cli_equals_one_reference_process_per_gap(tmp_path):
    """`Gap2Seq-core -gap-list` = the wrapper's loop of single-gap runs (src/Gap2Seq.py:132-220) in one process.  Every
    gap is compared with a run of the UNMODIFIED reference binary in single-gap mode on that gap's reads (when
    oracle/_ref is present), else with this tool's own single-gap mode - log line and sequence."""
    import os
    import subprocess
    import cases
    from gap2seq_b200.synth import write_reads_fasta
    ref = os.path.join(cases.ROOT, "oracle", "_ref", "bin", "Gap2Seq-core")
    exe = os.path.join(cases.ROOT, "gap2seq_b200", "lib", "Gap2Seq-core")
    single = ref if os.path.exists(ref) else exe
    w = make_workload("cfg1", 1.0)
    flank_l, flank_r = [45, 41, 60, 31, 52, 41], [41, 50, 44, 41, 38, 70]   # some longer than k + fuz, one without fuz
    lines, want_fa, want_log = [], [], []
    for gi, rec in enumerate(w.gaps[:6]):
        i = rec.seq.index("N")
        left, right = rec.seq[i - flank_l[gi]:i], rec.seq[i + rec.gap_len:i + rec.gap_len + flank_r[gi]]
        rp = str(tmp_path / f"reads{gi}.fa")
        write_reads_fasta(rp, w.reads_near([rec], reach=150))
        lines.append(f"gap{gi}\t{left}\t{right}\t{rec.gap_len}\t{rp}")
        wd = tmp_path / f"run{gi}"
        wd.mkdir()
        p = subprocess.run([single, "-k", "31", "-solid", "2", "-nb-cores", "1", "-randseed", "5", "-reads", rp, "-left", left, "-right", right,
                            "-length", str(rec.gap_len), "-filled", str(wd / "f.fa")], cwd=str(wd), capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
        want_fa.append((wd / "f.fa").read_text().splitlines()[1])
        want_log.append(cases.filter_log(p.stdout))
    (tmp_path / "gaps.tsv").write_text("\n".join(lines) + "\n")
    p = subprocess.run([exe, "-k", "31", "-solid", "2", "-randseed", "5", "-gap-list", str(tmp_path / "gaps.tsv"), "-filled", str(tmp_path / "out.fa")],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    got = (tmp_path / "out.fa").read_text().splitlines()
    assert got[0::2] == [f">gap{g}" for g in range(6)]
    assert got[1::2] == want_fa
    body = cases.filter_log(p.stdout)
    assert body.replace("Filled 6 gaps out of 6\n", "") == "".join(want_log)   # single-gap runs print no total
    assert sum(1 for s in want_fa if "N" not in s) >= 5
