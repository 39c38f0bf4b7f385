"""Seeded synthetic workloads for the gap-filling hot path.

The reference ships no sample data, so every parity test and bench.py build
their inputs here: a random genome (optionally with tandem / interspersed
repeats), simulated reads, and one-gap scaffold records of the form
``left flank + N*gap + right flank`` - the same shape the reference's
GapCutter emits (src/GapCutter.cpp:193-311) and that ``Gap2Seq-core
-scaffolds`` scans (src/Gap2Seq.cpp:340-420).  Headers contain a space
because GapCutter asserts one (src/GapCutter.cpp:159-160).

Pure numpy; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import bisect
import os
from dataclasses import dataclass, field

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[_a] = _b


def random_genome(n: int, rng: np.random.Generator) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def mutate(seq: np.ndarray, rate: float, rng: np.random.Generator) -> np.ndarray:
    """Substitute each base with probability `rate` by one of the 3 other bases."""
    out = seq.copy()
    if rate <= 0:
        return out
    hit = np.flatnonzero(rng.random(out.shape[0]) < rate)
    if hit.size:
        code = np.searchsorted(_ACGT, out[hit])  # A,C,G,T are sorted ASCII
        out[hit] = _ACGT[(code + rng.integers(1, 4, size=hit.size)) % 4]
    return out


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def add_repeats(genome: np.ndarray, rng: np.random.Generator, *, n_tandem: int, n_families: int,
                copies: tuple[int, int], family_len: tuple[int, int], divergence: tuple[float, float],
                inverted_copies: bool) -> list[int]:
    """Overwrite parts of `genome` in place with tandem arrays and an interspersed repeat family.

    Returns the start positions of the interspersed copies (gaps are placed next to them).
    `inverted_copies` inserts half of the copies reverse-complemented; this is what makes
    the reference's border sets order dependent, so it is an explicit switch.
    """
    n = genome.shape[0]
    starts: list[int] = []
    for _ in range(n_tandem):
        unit = random_genome(int(rng.integers(20, 61)), rng)
        reps = int(rng.integers(10, 41))
        pos = int(rng.integers(1000, n - unit.size * reps - 1000))
        for r in range(reps):
            genome[pos + r * unit.size: pos + (r + 1) * unit.size] = mutate(unit, 0.02, rng)
    for _ in range(n_families):
        cons = random_genome(int(rng.integers(family_len[0], family_len[1] + 1)), rng)
        ncopy = int(rng.integers(copies[0], copies[1] + 1))
        for _c in range(ncopy):
            pos = int(rng.integers(2000, n - cons.size - 2000))
            cp = mutate(cons, float(rng.uniform(*divergence)), rng)
            if inverted_copies and rng.random() < 0.5:
                cp = revcomp(cp)
            genome[pos: pos + cons.size] = cp
            starts.append(pos)
    return starts


def simulate_reads(genome: np.ndarray, coverage: float, read_len: int, err: float,
                   rng: np.random.Generator) -> np.ndarray:
    """[n_reads, read_len] uint8 ASCII; uniform starts, random strand, substitution errors."""
    n = genome.shape[0]
    n_reads = int(round(coverage * n / read_len))
    starts = rng.integers(0, n - read_len + 1, size=n_reads)
    simulate_reads.last_starts = starts          # genome position of every read (Workload.reads_near)
    reads = np.empty((n_reads, read_len), dtype=np.uint8)
    ar = np.arange(read_len)[None, :]
    for lo in range(0, n_reads, 1 << 20):   # gathered in chunks: the index matrix of 30 M reads would take 24 GB
        hi = min(n_reads, lo + (1 << 20))
        reads[lo:hi] = genome[starts[lo:hi, None] + ar]
    flip = rng.random(n_reads) < 0.5
    for lo in range(0, n_reads, 1 << 20):
        hi = min(n_reads, lo + (1 << 20))
        f = flip[lo:hi]
        reads[lo:hi][f] = _COMP[reads[lo:hi][f][:, ::-1]]
    if err > 0:
        hit = rng.random(reads.shape) < err
        nh = int(hit.sum())
        if nh:
            code = np.searchsorted(_ACGT, reads[hit])
            reads[hit] = _ACGT[(code + rng.integers(1, 4, size=nh)) % 4]
    return reads


def write_reads_fasta(path: str, reads: np.ndarray) -> None:
    n, L = reads.shape
    width = max(1, len(str(max(n - 1, 1))))
    hdr = np.char.zfill(np.arange(n).astype(f"U{width}"), width)
    hb = np.frombuffer("".join(hdr).encode(), dtype=np.uint8).reshape(n, width)
    rec = np.empty((n, 2 + width + 1 + L + 1), dtype=np.uint8)
    rec[:, 0] = ord(">")
    rec[:, 1] = ord("r")
    rec[:, 2:2 + width] = hb
    rec[:, 2 + width] = ord("\n")
    rec[:, 3 + width:3 + width + L] = reads
    rec[:, -1] = ord("\n")
    rec.tofile(path)


@dataclass
class GapRecord:
    name: str          # FASTA comment (contains a space)
    seq: str           # left flank + N*gap + right flank
    truth: str         # the bases that were replaced by N
    pos: int           # position of the gap in the genome
    gap_len: int


@dataclass
class Workload:
    name: str
    k: int
    solid: int
    dist_error: int
    fuz: int
    genome: np.ndarray
    reads: np.ndarray
    gaps: list[GapRecord] = field(default_factory=list)
    best_only: bool = False
    note: str = ""
    read_starts: np.ndarray | None = None   # genome position of reads[i] (set by make_workload)

    def reads_near(self, gaps: list, reach: int = 300) -> np.ndarray:
        """The reads that overlap the DP window of one of `gaps`: [pos - reach, pos + gap_len + dist_error + reach).
        The forward pass of a gap walks at most gap_len + dist_error + 2 fuz levels from the left flank and the right
        pass stays left of the right flank, so on a genome without long repeats the solid k-mers any of these gaps can
        touch - and their counts - are the same as with all reads: a bounded sample of a large workload that fills to
        the same result gap by gap (used for the CPU reference arm of bench.py; checked in tests/test_synth_sample.py)."""
        if self.read_starts is None:
            raise ValueError("workload without read positions")
        n, L = self.genome.shape[0], self.reads.shape[1]
        cover = np.zeros(n + 1, dtype=np.int32)
        for g in gaps:
            lo, hi = max(0, g.pos - reach), min(n, g.pos + g.gap_len + self.dist_error + reach)
            cover[lo] += 1
            cover[hi] -= 1
        inside = np.cumsum(cover[:-1]) > 0
        # a read [s, s + L) overlaps a window iff any of its positions is inside one
        csum = np.concatenate([[0], np.cumsum(inside, dtype=np.int64)])
        s0 = self.read_starts
        keep = (csum[np.minimum(s0 + L, n)] - csum[s0]) > 0
        return self.reads[keep]

    def write(self, outdir: str) -> dict:
        os.makedirs(outdir, exist_ok=True)
        reads = os.path.join(outdir, "reads.fa")
        gaps = os.path.join(outdir, "gaps.fa")
        write_reads_fasta(reads, self.reads)
        with open(gaps, "w") as f:
            for g in self.gaps:
                f.write(f">{g.name}\n{g.seq}\n")
        return {"reads": reads, "gaps": gaps}

    def read_strings(self) -> list[str]:
        return [r.tobytes().decode() for r in self.reads]


def place_gaps(genome: np.ndarray, lengths: np.ndarray, rng: np.random.Generator, *, flank: int,
               near: list[int] | None = None, min_sep: int = 200, est_jitter: float = 0.0) -> list[GapRecord]:
    """One-gap records.  `near`: optional anchor positions, each gap starts <=150 bp upstream of one."""
    n = genome.shape[0]
    # accepted windows, sorted by start.  They are pairwise separated, hence sorted by end as well, so a candidate
    # [a, b) conflicts with some window iff it conflicts with the last window that starts before b + min_sep
    # (the same accept / reject decisions as testing every window, in O(log n)).
    starts: list[int] = []
    ends: list[int] = []
    out: list[GapRecord] = []
    for gi, L in enumerate(int(x) for x in lengths):
        for _try in range(1000):
            if near:
                a = near[int(rng.integers(0, len(near)))]
                p = a - int(rng.integers(0, 151))
            else:
                p = int(rng.integers(flank + 1, n - L - flank - 1))
            if p - flank < 0 or p + L + flank > n:
                continue
            j = bisect.bisect_left(starts, p + L + flank + min_sep)   # windows [0, j) start before b + min_sep
            if j == 0 or ends[j - 1] + min_sep <= p - flank:
                break
        else:
            continue
        j = bisect.bisect_left(starts, p - flank)
        starts.insert(j, p - flank)
        ends.insert(j, p + L + flank)
        est = L if est_jitter <= 0 else max(1, int(round(L * (1 + rng.uniform(-est_jitter, est_jitter)))))
        left = genome[p - flank:p].tobytes().decode()
        right = genome[p + L:p + L + flank].tobytes().decode()
        out.append(GapRecord(name=f"gap{gi} genome:{p}-{p + L}", seq=left + "N" * est + right,
                             truth=genome[p:p + L].tobytes().decode(), pos=p, gap_len=est))
    return out


def _loguniform(rng, lo, hi, n):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.int64)


def make_workload(name: str, scale: float = 1.0, *, inverted_copies: bool = False, seed: int | None = None,
                  with_reads: bool = True, repeat_load: float = 1.0) -> Workload:
    """The five named configurations of BASELINE.json; `scale` shrinks genome and gap count
    for tests (scale=1.0 is the full configuration).  with_reads=False skips the read simulation (the genome is the
    same, the gap list is NOT: ranks that receive the graph by broadcast place their own gaps, bench.py)."""
    if not with_reads:
        global simulate_reads
        real = simulate_reads
        simulate_reads = lambda g, cov, rl, err, rng: np.zeros((0, rl), dtype=np.uint8)   # noqa: E731
        try:
            return make_workload(name, scale, inverted_copies=inverted_copies, seed=seed, repeat_load=repeat_load)
        finally:
            simulate_reads = real
    flank = 100
    w = _make_workload(name, scale, inverted_copies, seed, repeat_load, flank)
    w.read_starts = getattr(simulate_reads, "last_starts", None)
    return w


def _make_workload(name, scale, inverted_copies, seed, repeat_load, flank) -> Workload:
    if name == "cfg1":      # 100 kbp, 30x error-free 100 bp reads, 20 gaps of 50-500 bp
        rng = np.random.default_rng(1 if seed is None else seed)
        n = max(20000, int(100_000 * scale)); ngap = max(3, int(round(20 * scale)))
        g = random_genome(n, rng)
        reads = simulate_reads(g, 30, 100, 0.0, rng)
        gaps = place_gaps(g, rng.integers(50, 501, size=ngap), rng, flank=flank)
        return Workload(name, 31, 2, 500, 10, g, reads, gaps)
    if name == "cfg2":      # S. aureus scale: 2.9 Mbp, 50x 101 bp reads, 1 % substitutions, ~170 gaps <= 3 kbp
        rng = np.random.default_rng(2 if seed is None else seed)
        n = max(30000, int(2_900_000 * scale)); ngap = max(4, int(round(170 * scale)))
        g = random_genome(n, rng)
        reads = simulate_reads(g, 50, 101, 0.01, rng)
        gaps = place_gaps(g, _loguniform(rng, 50, 3000, ngap), rng, flank=flank)
        return Workload(name, 31, 2, 500, 10, g, reads, gaps)
    if name == "cfg3":      # repeat-rich 5 Mbp: tandem arrays + interspersed family, all-paths
        rng = np.random.default_rng(3 if seed is None else seed)
        n = max(60000, int(5_000_000 * scale)); ngap = max(4, int(round(60 * scale)))
        g = random_genome(n, rng)
        # repeat_load (cfg3 only, default 1.0 = the configuration the goldens were made with): multiplies the number of
        # interspersed copies, i.e. the fraction of the genome they cover (1.6 % at 1.0).  A walk that leaves one copy
        # then runs into the next one before the DP ends and branches again: at ~20 % coverage (repeat_load 12) a gap
        # holds 10^6 - 10^7 DP cells, the regime of the reference probe in SURVEY section 6
        rl = float(repeat_load)
        starts = add_repeats(g, rng, n_tandem=max(2, int(40 * scale)), n_families=1,
                             copies=(max(8, int(100 * scale * rl)), max(12, int(300 * scale * rl))),
                             family_len=(300, 500), divergence=(0.02, 0.05), inverted_copies=inverted_copies)
        reads = simulate_reads(g, 30, 100, 0.0, rng)
        gaps = place_gaps(g, rng.integers(100, 701, size=ngap), rng, flank=flank, near=starts)
        return Workload(name, 31, 2, 500, 10, g, reads, gaps,
                        note=f"inverted_copies={inverted_copies} repeat_load={rl:g}")
    if name == "cfg4":      # chromosome scale: 100 Mbp, 30x, 10k gaps 100 bp - 5 kbp
        rng = np.random.default_rng(4 if seed is None else seed)
        n = max(40000, int(100_000_000 * scale)); ngap = max(4, int(round(10_000 * scale)))
        g = random_genome(n, rng)
        reads = simulate_reads(g, 30, 100, 0.0, rng)
        gaps = place_gaps(g, _loguniform(rng, 100, 5000, ngap), rng, flank=flank)
        return Workload(name, 31, 2, 500, 10, g, reads, gaps)
    if name == "cfg5":      # insertion genotyping: reference + insertions, 40x reads from the donor
        rng = np.random.default_rng(5 if seed is None else seed)
        n = max(40000, int(20_000_000 * scale)); nins = max(4, int(round(5000 * scale)))
        ref = random_genome(n, rng)
        pos = np.sort(rng.choice(np.arange(500, n - 500, 2600), size=min(nins, (n - 1000) // 2600), replace=False))
        lens = rng.integers(50, 2001, size=pos.size)
        parts, gaps, prev = [], [], 0
        for i, (p, L) in enumerate(zip(pos.tolist(), lens.tolist())):
            ins = random_genome(L, rng)
            parts += [ref[prev:p], ins]
            prev = p
            left = ref[p - 41:p].tobytes().decode(); right = ref[p:p + 41].tobytes().decode()
            # gap record = ref flank + N*len(insert) + ref flank (the wrapper's cut_vcf is broken)
            gaps.append(GapRecord(name=f"ins{i} ref:{p} len={L}", seq=left + "N" * L + right,
                                  truth=ins.tobytes().decode(), pos=p, gap_len=L))
        parts.append(ref[prev:])
        donor = np.concatenate(parts)
        reads = simulate_reads(donor, 40, 100, 0.0, rng)
        return Workload(name, 31, 2, 500, 10, donor, reads, gaps, note="-v/-R genotyping as gap records")
    raise ValueError(f"unknown workload {name!r}")


def make_order_dependent_workload(seed: int = 77, n_plain: int = 3) -> Workload:
    """A workload with ONE gap whose border provably receives both strands of a k-mer on one level, between
    `n_plain` ordinary gaps before and after it.

    Two copies of a 60 bp segment L are followed by equally long but different segments a / b and then by a 101 bp
    segment S in copy 1 and by revcomp(S) in copy 2.  A walk that starts in L branches after it; 35 bases into S (its
    middle k-mer, k = 31) the walk through copy 1 stands on k-mer K and the walk through copy 2 on revcomp(K) at the same
    level: the reference inserts both strands of one node into one unordered_set (src/Gap2Seq.cpp:1064), its result
    depends on hashtable order, and the engine must flag the gap G2S_ST_ORDER_DEPENDENT.  The gap starts right after
    L of copy 1."""
    rng = np.random.default_rng(seed)
    g = random_genome(40_000, rng)
    L, a, b, S = (random_genome(n, rng) for n in (60, 50, 50, 101))
    p1, p2 = 12_000, 25_000
    c1 = np.concatenate([L, a, S])
    c2 = np.concatenate([L, b, revcomp(S)])
    g[p1:p1 + c1.size] = c1
    g[p2:p2 + c2.size] = c2
    reads = simulate_reads(g, 30, 100, 0.0, rng)
    plain = place_gaps(np.concatenate([g[:p1 - 500], g[p2 + 1000:]]), rng.integers(80, 400, size=2 * n_plain), rng, flank=100)
    gs, glen = p1 + L.size, 320
    od = GapRecord(name=f"hairpin genome:{gs}-{gs + glen}", seq=g[gs - 100:gs].tobytes().decode() + "N" * glen +
                   g[gs + glen:gs + glen + 100].tobytes().decode(), truth=g[gs:gs + glen].tobytes().decode(), pos=gs, gap_len=glen)
    return Workload("order_dependent", 31, 2, 500, 10, g, reads, plain[:n_plain] + [od] + plain[n_plain:],
                    note="one Q3 gap between plain gaps")
