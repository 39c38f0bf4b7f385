"""The insertion genotyping front-end (gap2seq_b200/vcf.py, SURVEY 8f rank 4 / BASELINE configs[4]): VCF + reference ->
gap records that fill to the inserted sequences."""
import numpy as np

import emul_lib
import oracle_lib
from gap2seq_b200 import synth, vcf


def test_vcf_records_equal_the_directly_built_cfg5_records_and_fill_to_the_insertions(tmp_path):
    rng = np.random.default_rng(55)
    n = 60000
    ref = synth.random_genome(n, rng)
    refs = ref.tobytes().decode()
    pos = list(range(2000, n - 3000, 5200))
    ins = [synth.random_genome(int(rng.integers(50, 600)), rng).tobytes().decode() for _ in pos]
    # donor = reference with the insertions; reads come from the donor
    parts, prev = [], 0
    for p, s in zip(pos, ins):
        parts += [refs[prev:p], s]
        prev = p
    donor = np.frombuffer(("".join(parts) + refs[prev:]).encode(), dtype=np.uint8)
    reads = synth.simulate_reads(donor, 40, 100, 0.0, rng)
    fa, vc = str(tmp_path / "ref.fa"), str(tmp_path / "ins.vcf")
    with open(fa, "w") as f:
        f.write(">chr1 synthetic reference\n")
        for j in range(0, n, 80):
            f.write(refs[j:j + 80] + "\n")
    vcf.write_vcf(vc, "chr1", refs, list(zip(pos, ins)))
    names, records, short = vcf.cut_vcf(vc, fa, 31, 10, str(tmp_path / "gaps.fa"), str(tmp_path / "gaps.bed"))
    assert len(records) == len(pos) and not any(short)
    for p, s, r in zip(pos, ins, records):   # the geometry synth.make_workload("cfg5") uses
        assert r == refs[p - 41:p] + "N" * len(s) + refs[p:p + 41]
    assert open(tmp_path / "gaps.bed").read().splitlines()[0] == f"chr1\t{pos[0] - 41}\t{pos[0] + 41}"
    assert vcf.main(["-v", vc, "-R", fa, "-g", str(tmp_path / "g2.fa")]) == 0
    assert open(tmp_path / "g2.fa").read() == open(tmp_path / "gaps.fa").read()
    # fill them: every insertion comes back, base for base
    g = oracle_lib.OracleGraph.from_reads(31, 2, reads)
    eng = emul_lib.EmulEngine(31, g.kmers())
    oracle_lib.srand(1)
    filled, log, gc, fc, _ = eng.fill_records(records, names, dist_error=100, fuz=10)
    assert gc == len(pos)
    ok = 0
    for p, s, f in zip(pos, ins, filled):
        if "N" in f:
            continue   # a coverage hole (40x leaves very few)
        assert f.upper() == refs[p - 41:p] + s + refs[p:p + 41]
        ok += 1
    assert ok >= len(pos) - 1 and fc == ok


def test_symbolic_insertions_and_other_variants(tmp_path):
    fa, vc = str(tmp_path / "r.fa"), str(tmp_path / "v.vcf")
    seq = "ACGT" * 50
    open(fa, "w").write(">c1 x\n" + seq + "\n")
    open(vc, "w").write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
                        "c1\t100\t.\tT\t<INS>\t.\t.\tSVTYPE=INS;SVLEN=25\n"      # symbolic: length from SVLEN
                        "c1\t120\t.\tT\tG\t.\t.\t.\n"                           # SNV: ignored
                        "c1\t130\t.\tCG\tC\t.\t.\t.\n"                          # deletion: ignored
                        "c1\t5\t.\tA\tATT\t.\t.\t.\n"                           # left flank shorter than k + fuz: reported
                        "c1\t198\t.\tC\tCAAAA\t.\t.\t.\n")                      # right flank shorter than k: dropped
    names, records, short = vcf.cut_vcf(vc, fa, 31, 10)
    assert [r.count("N") for r in records] == [25, 2]
    assert records[0] == seq[100 - 41:100] + "N" * 25 + seq[100:141]
    assert short == [False, True]
