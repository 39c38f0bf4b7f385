"""The wide level step of the forward pass (wide_forward_level in gap2seq_b200/csrc/g2s_fill.cuh: per-level de-duplication
in a shared-memory hash table, hash partitions, partition splitting) on the CPU through the host emulation, against the
oracle.  The table size is a run-time parameter there: 8192 entries is the device's size (one partition for most
levels), 64 and 16 force partitions and splits on the same gaps, 0 switches the step off (general path only) - all four
must give the oracle's results on repeat-rich gaps (wide borders, saturated counts, general subgraphs)."""
import pytest

import emul_lib
import oracle_lib
import parity
from gap2seq_b200 import synth
from gap2seq_b200.scaffold import record_gaps

MODES = {"all_paths": dict(skip_confident=False, all_paths=True), "best_only": dict(skip_confident=False, all_paths=False)}


@pytest.fixture(scope="module")
def repeat_workloads():
    out = []
    for name, scale, inv in (("cfg3", 0.02, False), ("cfg3", 0.015, True), ("cfg2", 0.05, False)):
        w = synth.make_workload(name, scale, inverted_copies=inv)
        og = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
        out.append((w, og, emul_lib.EmulEngine(w.k, og.kmers())))
    return out


@pytest.mark.parametrize("cap", [8192, 64, 16, 0])
@pytest.mark.parametrize("mode", list(MODES))
def test_wide_level_step_matches_oracle(cap, mode, repeat_workloads, monkeypatch):
    monkeypatch.setenv("G2S_EMUL_WIDE_CAP", str(cap))
    monkeypatch.setenv("G2S_NO_CHAIN", "1")   # read once per process by the emulation; harmless either way
    total = 0
    for w, og, em in repeat_workloads:
        specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error, **MODES[mode])]
        out = parity.compare_with_oracle(em.fill_batch, og, specs)
        assert out["compared"] + len(out["order_dependent"]) == len(specs)
        total += out["compared"]
    assert total > 0
