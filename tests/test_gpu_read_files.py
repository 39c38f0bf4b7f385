"""Reads files parsed on the device (g2s_build_from_read_files): FASTA with sequences over one or several lines, CR LF
line ends, four-line FASTQ, several files - the table must equal the one counted from the same reads handed over as
arrays (g2s_build_from_reads), which the parity tests check against the oracle."""
import numpy as np
import pytest

import cases
from gap2seq_b200 import engine

pytestmark = pytest.mark.gpu


def _tables_equal(a, b):
    ka, aa = a.export_kmers()
    kb, ab = b.export_kmers()
    return np.array_equal(ka, kb) and np.array_equal(aa, ab)


def _write(path, reads, fmt, wrap=0, crlf=False, lower_every=0):
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        for i, r in enumerate(reads):
            s = r.tobytes().decode()
            if lower_every and i % lower_every == 0:
                s = s.lower()
            if fmt == "fasta":
                f.write(f">read{i} some comment{nl}")
                if wrap:
                    for j in range(0, len(s), wrap):
                        f.write(s[j:j + wrap] + nl)
                else:
                    f.write(s + nl)
            else:
                f.write(f"@read{i}{nl}{s}{nl}+{nl}{'I' * len(s)}{nl}")


@pytest.mark.parametrize("fmt,wrap,crlf", [("fasta", 0, False), ("fasta", 37, False), ("fasta", 60, True), ("fastq", 0, False), ("fastq", 0, True)])
def test_device_parsed_reads_give_the_same_table(tmp_path, fmt, wrap, crlf):
    w = cases.build(cases.CASE_BY_NAME["cfg1_small"])
    p = str(tmp_path / ("reads." + ("fa" if fmt == "fasta" else "fq")))
    _write(p, w.reads, fmt, wrap, crlf)
    with engine.Engine(k=w.k) as a, engine.Engine(k=w.k) as b:
        assert a.build_from_read_files([p], w.solid) == w.reads.shape[0]
        b.build_from_reads(w.reads, w.solid)
        assert a.num_kmers() == b.num_kmers() > 0
        assert _tables_equal(a, b)


def test_several_files_short_reads_and_invalid_characters(tmp_path):
    w = cases.build(cases.CASE_BY_NAME["cfg1_small"])
    reads = w.reads.copy()
    reads[5, 40] = ord("N")                      # k-mers over an N are not counted (Data.hpp:176)
    half = reads.shape[0] // 2
    p1, p2 = str(tmp_path / "a.fa"), str(tmp_path / "b.fa")
    _write(p1, reads[:half], "fasta", 50)
    _write(p2, reads[half:], "fasta", 0)
    with open(p2, "a") as f:                     # a read shorter than k, an empty record, no line break at the end
        last = "".join("ACGT"[x] for x in np.random.default_rng(5).integers(0, 4, 40))   # 10 k-mers seen once: not solid
        f.write(">short\nACGTACGT\n>empty\n>last\n" + last)
    with engine.Engine(k=w.k) as a, engine.Engine(k=w.k) as b:
        assert a.build_from_read_files([p1, p2], w.solid) == reads.shape[0] + 3
        b.build_from_reads(reads, w.solid)
        assert _tables_equal(a, b)
        with pytest.raises(engine.G2SError, match="neither FASTA"):
            bad = tmp_path / "x.txt"
            bad.write_text("hello\nworld\n")
            a.build_from_read_files([str(bad)], 2)
        with pytest.raises(engine.G2SError, match="cannot open"):
            a.build_from_read_files([str(tmp_path / "missing.fa")], 2)
