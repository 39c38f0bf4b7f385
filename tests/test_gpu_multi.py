"""Several GPUs behind ONE engine (g2s_create_multi; the reference hands records to -nb-cores Dispatcher
workers, src/Gap2Seq.cpp:296-323, 426-432): every wave of gaps is split longest-processing-time-first over the engines,
each runs its shard on its own host thread, the per-gap results are merged in input order and ONE libc rand() stream
drives the tracebacks - so the output must not depend on the number of engines.

On a one-GPU box the same code path runs with several engines on the one GPU (G2S_ALLOW_DUPLICATE_DEVICES=1: the table is
replicated by a device-to-device copy instead of ncclBroadcast); with two or more GPUs the real layout runs too."""
import ctypes
import os

import pytest

import oracle_lib
from gap2seq_b200 import engine, synth
from gap2seq_b200.scaffold import record_gaps

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def run_records(eng, w, seed=1):
    ctypes.CDLL(None).srand(seed)
    return eng.fill_records([g.seq for g in w.gaps], [g.name for g in w.gaps], dist_error=w.dist_error, fuz=w.fuz,
                            all_paths=not w.best_only)


@pytest.fixture(scope="module")
def one_gpu_outputs():
    out = {}
    for name, scale in (("cfg2", 0.5), ("cfg3", 0.3), ("cfg4", 0.02)):
        w = synth.make_workload(name, scale)
        e = engine.Engine(k=w.k, max_fuz=w.fuz)
        e.build_from_reads(w.reads, w.solid)
        out[name] = (w, run_records(e, w))
        e.close()
    return out


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
@pytest.mark.parametrize("n_eng", [2, 3])
def test_engines_sharing_one_gpu_give_the_one_engine_output(name, n_eng, one_gpu_outputs, monkeypatch):
    monkeypatch.setenv("G2S_ALLOW_DUPLICATE_DEVICES", "1")
    w, ref = one_gpu_outputs[name]
    e = engine.Engine(k=w.k, max_fuz=w.fuz, devices=[0] * n_eng)
    try:
        assert e.L.g2s_num_devices(e.h) == n_eng
        e.build_from_reads(w.reads, w.solid)
        got = run_records(e, w)
        assert got[0] == ref[0] and got[1] == ref[1] and got[2:4] == ref[2:4]
        # the batch call as well: same gaps, same rand() stream
        specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error, all_paths=not w.best_only)]
        ctypes.CDLL(None).srand(1)
        res, fills = e.fill_batch(specs)
        assert [r.count for r in res] == [r.count for r in ref[4]]
        assert [r.rand_calls for r in res] == [r.rand_calls for r in ref[4]]
    finally:
        e.close()


def test_duplicate_devices_are_refused_without_the_test_switch(monkeypatch):
    monkeypatch.delenv("G2S_ALLOW_DUPLICATE_DEVICES", raising=False)
    with pytest.raises(engine.G2SError):
        engine.Engine(k=31, max_fuz=10, devices=[0, 0])


@pytest.mark.parametrize("name", ["cfg2", "cfg4"])
def test_all_gpus_of_the_box_give_the_one_gpu_output(name, one_gpu_outputs):
    n = n_gpus()
    if n < 2:
        pytest.skip("one GPU on this box")
    w, ref = one_gpu_outputs[name]
    e = engine.Engine(k=w.k, max_fuz=w.fuz, devices=list(range(n)))
    try:
        e.build_from_reads(w.reads, w.solid)
        got = run_records(e, w)
        assert got[0] == ref[0] and got[1] == ref[1] and got[2:4] == ref[2:4]
    finally:
        e.close()


def test_multi_engine_output_matches_the_oracle(one_gpu_outputs, monkeypatch):
    monkeypatch.setenv("G2S_ALLOW_DUPLICATE_DEVICES", "1")
    w, _ = one_gpu_outputs["cfg4"]
    e = engine.Engine(k=w.k, max_fuz=w.fuz, devices=[0, 0])
    try:
        e.build_from_reads(w.reads, w.solid)
        keys, _ = e.export_kmers()
        og = oracle_lib.OracleGraph.from_kmers(w.k, keys)
        got = run_records(e, w)
        oracle_lib.srand(1)
        for rec, f in zip(w.gaps, got[0]):
            assert og.fill_record(rec.seq, rec.name, d_err=w.dist_error, fuz=w.fuz)[0] == f
    finally:
        e.close()
