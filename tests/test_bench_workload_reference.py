"""The bench workload itself (BASELINE.json configs[1], all 170 gaps at full size) through the product's gap driver and the
host emulation of the device algorithm, against the UNMODIFIED reference binary run on the same files: filled FASTA and
log lines byte for byte.  (The CUDA kernels are compared with the oracle on the same 170 gaps by the -m gpu tests and
tools/gpu_check_oracle.py; this closes the loop to the reference itself at full size.)"""
import os
import subprocess

import pytest

import cases
import emul_lib
import oracle_lib
from gap2seq_b200 import synth

REF = os.path.join(cases.ROOT, "oracle", "_ref", "bin", "Gap2Seq-core")


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built (oracle/build_ref.sh)")
def test_cfg2_full_size_equals_the_reference_binary(tmp_path):
    w = synth.make_workload("cfg2", 1.0)
    paths = w.write(str(tmp_path))
    out = str(tmp_path / "ref_filled.fa")
    p = subprocess.run([REF, "-k", "31", "-solid", "2", "-fuz", "10", "-dist-error", "500", "-randseed", "1", "-nb-cores", "1",
                        "-reads", paths["reads"], "-scaffolds", paths["gaps"], "-filled", out],
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    g = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
    eng = emul_lib.EmulEngine(w.k, g.kmers())
    oracle_lib.srand(1)
    filled, log, gc, fc, res = eng.fill_records([x.seq for x in w.gaps], [x.name for x in w.gaps], dist_error=500, fuz=10)
    assert not any(r.status & 1 for r in res)
    assert "".join(f">{x.name}\n{f}\n" for x, f in zip(w.gaps, filled)) == open(out).read()
    assert log + f"Filled {fc} gaps out of {gc}\n" == cases.filter_log(p.stdout)
    assert (gc, fc) == (170, 170)
