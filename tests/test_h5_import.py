"""the solid k-mers of a GATB graph file (<reads>.h5, group dsk/solid) are read straight from the
HDF5 container by gap2seq_b200/csrc/g2s_h5.h (no GATB, no libhdf5).  Fixture: tests/golden/cfg1_small.h5.gz, the file
the unmodified reference wrote for the cfg1_small case (tests/golden/make_golden.py h5)."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import cases
import oracle_lib
from gap2seq_b200 import engine

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "cfg1_small.h5.gz")
REF = os.path.join(cases.ROOT, "oracle", "_ref", "bin", "Gap2Seq-core")


def _fixture(tmp_path) -> str:
    p = tmp_path / "reads.fa.h5"
    p.write_bytes(gzip.decompress(open(FIXTURE, "rb").read()))
    return str(p)


def test_solid_kmers_of_the_reference_graph_file_equal_the_oracle_graph(tmp_path):
    kmers, ksize = engine.h5_solid_kmers(_fixture(tmp_path))
    assert ksize == 31
    w = cases.build(cases.CASE_BY_NAME["cfg1_small"])
    want = np.sort(oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads).kmers())
    assert kmers.size == want.size and np.unique(kmers).size == kmers.size   # partitions are disjoint
    assert np.array_equal(np.sort(kmers), want)


def test_compressed_graph_file_written_by_gatb(tmp_path):
    """GATB's own -out-compress (shuffle + deflate on every dataset, CollectionHDF5Patch.hpp:184-187): the fixture
    tests/golden/cfg1_small_compressed.h5.gz was written by GATB's Graph::create with that option
    (oracle/make_h5_compressed_fixture.sh); the node set must be the same as that of the plain file."""
    p = tmp_path / "c.h5"
    p.write_bytes(gzip.decompress(open(os.path.join(HERE, "golden", "cfg1_small_compressed.h5.gz"), "rb").read()))
    kc, ksize = engine.h5_solid_kmers(str(p))
    kp, _ = engine.h5_solid_kmers(_fixture(tmp_path))
    assert ksize == 31 and kc.size == kp.size == 19955
    assert np.array_equal(np.sort(kc), np.sort(kp))


def test_reader_rejects_what_it_does_not_understand(tmp_path):
    bad = tmp_path / "x.h5"
    bad.write_bytes(b"not an hdf5 file at all" * 10)
    with pytest.raises(engine.G2SError, match="not an HDF5 file"):
        engine.h5_solid_kmers(str(bad))
    raw = bytearray(gzip.decompress(open(FIXTURE, "rb").read()))
    raw[8] = 2                                   # a version-2 superblock is a different layout: refused, not guessed at
    bad.write_bytes(bytes(raw))
    with pytest.raises(engine.G2SError, match="superblock version 2"):
        engine.h5_solid_kmers(str(bad))
    bad.write_bytes(bytes(raw[:70000]))          # truncated file: every address is bounds-checked
    with pytest.raises(engine.G2SError):
        engine.h5_solid_kmers(str(bad))
    with pytest.raises(engine.G2SError, match="cannot open"):
        engine.h5_solid_kmers(str(tmp_path / "missing.h5"))


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref is not built")
def test_graph_file_written_now_by_the_reference(tmp_path):
    """A second case, written by the reference binary during the test (another k, other reads)."""
    w = cases.build(cases.CASE_BY_NAME["cfg1_small"])
    paths = w.write(str(tmp_path))
    cmd = [REF, "-k", "25", "-solid", "3", "-nb-cores", "1", "-randseed", "1", "-reads", paths["reads"], "-scaffolds", paths["gaps"],
           "-filled", str(tmp_path / "filled.fa")]
    subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, check=True, timeout=600)
    kmers, ksize = engine.h5_solid_kmers(str(tmp_path / ".h5"))
    assert ksize == 25
    want = np.sort(oracle_lib.OracleGraph.from_reads(25, 3, w.reads).kmers())
    assert np.array_equal(np.sort(kmers), want)


@pytest.mark.gpu
def test_import_h5_builds_the_same_table_as_counting_the_reads(tmp_path):
    w = cases.build(cases.CASE_BY_NAME["cfg1_small"])
    with engine.Engine(k=w.k) as a, engine.Engine(k=w.k) as b:
        a.import_h5(_fixture(tmp_path))
        b.build_from_reads(w.reads, w.solid)
        ka, aa = a.export_kmers()
        kb, ab = b.export_kmers()
        assert np.array_equal(ka, kb) and np.array_equal(aa, ab)
    with engine.Engine(k=27) as c, pytest.raises(engine.G2SError, match="31-mers"):
        c.import_h5(_fixture(tmp_path))


@pytest.mark.gpu
def test_cli_loads_the_graph_file_next_to_the_reads_like_the_reference(tmp_path):
    """Gap2Seq::execute loads <reads>.h5 when it is readable (src/Gap2Seq.cpp:171, 195-197) - so does the clone; the
    reads file itself is then never opened (here it does not even exist)."""
    name = "cfg1_small"
    w = cases.build(cases.CASE_BY_NAME[name])
    paths = w.write(str(tmp_path))
    os.remove(paths["reads"])
    os.replace(_fixture(tmp_path), paths["reads"] + ".h5") if _fixture(tmp_path) != paths["reads"] + ".h5" else None
    exe = os.path.join(cases.ROOT, "gap2seq_b200", "lib", "Gap2Seq-core")
    out = tmp_path / "filled.fa"
    p = subprocess.run([exe] + cases.ref_cmd_flags(cases.CASE_BY_NAME[name]) + ["-reads", paths["reads"], "-scaffolds", paths["gaps"], "-filled", str(out)],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "Loading from " + paths["reads"] + ".h5" in p.stdout
    fa, _ = cases.golden_paths(name)
    assert out.read_text() == open(fa).read()
