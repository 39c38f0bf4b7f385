"""The chain tier on the device (k_chain_walk + k_chain_finish, gap2seq_b200/csrc/g2s_chain.cu): gaps in clean unique
sequence end there; its results must equal both the oracle's and those of the tiers with tables (the same engine with
G2S_NO_CHAIN=1), field by field, and gaps it cannot prove to be chains must go on to the thin tier."""
import ctypes

import numpy as np
import pytest

import oracle_lib
import parity
from gap2seq_b200 import engine, synth
from gap2seq_b200.scaffold import record_gaps

pytestmark = pytest.mark.gpu

FIELDS = ("count", "left_fuz", "right_fuz", "status", "fill_len", "substats_valid", "forward_count", "n_path_lengths",
          "reached_fuz", "rand_calls", "cells_right", "cells_left")


def build(name, scale, monkeypatch=None, no_chain=False):
    w = synth.make_workload(name, scale)
    if monkeypatch is not None:
        if no_chain:
            monkeypatch.setenv("G2S_NO_CHAIN", "1")
        else:
            monkeypatch.delenv("G2S_NO_CHAIN", raising=False)
    e = engine.Engine(k=w.k, max_fuz=w.fuz)
    e.build_from_reads(w.reads, w.solid)
    return w, e


@pytest.mark.parametrize("name,scale", [("cfg1", 1.0), ("cfg4", 0.05), ("cfg5", 0.1)])
@pytest.mark.parametrize("mode", ["all_paths", "best_only", "all_upper"])
def test_clean_workloads_end_in_the_chain_tier_and_equal_the_table_tiers(name, scale, mode, monkeypatch):
    kw = {"all_paths": dict(skip_confident=False, all_paths=True), "best_only": dict(skip_confident=False, all_paths=False),
          "all_upper": dict(skip_confident=True, all_paths=True)}[mode]
    w, e = build(name, scale, monkeypatch, no_chain=False)
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error, **kw)]
    ctypes.CDLL(None).srand(5)
    res, fills = e.fill_batch(specs)
    t = e.timing()
    assert t.gaps_chain == len(specs) and t.launches_chain == 2 and t.launches_thin == 0
    assert t.cells_chain == sum(r.cells_left + r.cells_right for r in res)
    keys, _ = e.export_kmers()
    e.close()
    w2, e2 = build(name, scale, monkeypatch, no_chain=True)
    ctypes.CDLL(None).srand(5)
    res2, fills2 = e2.fill_batch(specs)
    t2 = e2.timing()
    assert t2.gaps_chain == 0 and t2.launches_thin >= 1
    assert fills == fills2
    for a, b in zip(res, res2):
        assert [getattr(a, f) for f in FIELDS] == [getattr(b, f) for f in FIELDS]
        assert a.substats.astuple() == b.substats.astuple()
    og = oracle_lib.OracleGraph.from_kmers(w.k, keys)
    out = parity.compare_with_oracle(e2.fill_batch, og, specs[:300])
    assert out["compared"] == min(300, len(specs))
    e2.close()


def test_gaps_with_bubbles_go_on_to_the_thin_tier():
    w, e = build("cfg2", 0.3)
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error)]
    keys, _ = e.export_kmers()
    og = oracle_lib.OracleGraph.from_kmers(w.k, keys)
    out = parity.compare_with_oracle(e.fill_batch, og, specs)
    t = e.timing()
    assert out["compared"] + len(out["order_dependent"]) == len(specs)
    assert t.launches_chain == 2 and t.launches_thin >= 1 and t.gaps_chain < len(specs)
    e.close()
