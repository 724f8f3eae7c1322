"""Seeded synthetic inputs of the BASELINE.json configurations 3, 4 and 5,
shared by bench.py (--config c3|c4|c5), tools/bench_configs.py and the full-size parity tests.
Like synth.py everything is block-keyed, so any rank can materialise its own shard (a range of
subject letters, a range of reads) without generating the rest.

  C3  countPDict of 1M 36-mers, PDict(tb.start=1, tb.end=12), max.mismatch=2, 100 Mbp subject
      (the SolexaYi2-benchmark shape, /root/reference/vignettes/SolexaYi2-benchmark.txt)
  C4  vcountPDict of 100k 25-mer probes against 10M reads of 150 bp (collapse=1 / 2)
  C5  2M patterns of 15-40 letters, PDict(tb.start=1, tb.width=15), fixed=FALSE, a 3.1 Gbp genome
      of 24 chromosomes with isolated IUPAC ambiguity letters at rate 1e-3
"""
import numpy as np

from . import synth

BASES = synth.BASES

# ---- C3 -------------------------------------------------------------------------------------

C3_WIDTH, C3_TB_END = 36, 12


def c3_dictionary(M=1_000_000, n=100_000_000):
    """M 36-mers, 10 % cut from the subject with 0-3 substitutions behind the Trusted Band."""
    return synth.dictionary(M, C3_WIDTH, n, 1, seed=4, subject_seed=5, nsub=3, sub_lo=C3_TB_END, with_n=False)


def c3_subject(lo, hi, n=100_000_000):
    return synth.subject_region(lo, hi, n, seed=5, with_n=False)


# ---- C4 -------------------------------------------------------------------------------------

C4_READ, C4_WIDTH = 150, 25
READ_BLOCK = 1 << 16


def c4_probes(M=100_000):
    return synth.dictionary(M, C4_WIDTH, 10_000_000, 1, seed=6, frac_planted=0.0)


def c4_reads(lo, hi, probes, frac_planted=0.05, seed=7):
    """Reads [lo, hi) of the set: i.i.d. letters, `frac_planted` of them carrying one probe at a
    random offset.  uint8 [hi - lo, 150]; read r belongs to block r // 65536."""
    out = np.empty((hi - lo, C4_READ), dtype=np.uint8)
    for b in range(lo // READ_BLOCK, (hi + READ_BLOCK - 1) // READ_BLOCK):
        g = np.random.Generator(np.random.PCG64([seed, b]))
        blk = BASES[g.integers(0, 4, size=(READ_BLOCK, C4_READ), dtype=np.uint8)]
        k = int(READ_BLOCK * frac_planted)
        rows = g.choice(READ_BLOCK, size=k, replace=False)
        which = g.integers(0, len(probes), size=k)
        off = g.integers(0, C4_READ - C4_WIDTH + 1, size=k)
        idx = off[:, None] + np.arange(C4_WIDTH)[None, :]
        blk[rows[:, None], idx] = probes[which]
        a, z = max(lo, b * READ_BLOCK), min(hi, (b + 1) * READ_BLOCK)
        out[a - lo:z - lo] = blk[a - b * READ_BLOCK:z - b * READ_BLOCK]
    return out


# ---- C5 -------------------------------------------------------------------------------------

C5_TB_WIDTH, C5_MIN_W, C5_MAX_W = 15, 15, 40
AMBIGUITY = np.array([3, 5, 9, 6, 10, 12, 7, 11, 13, 14, 15], dtype=np.uint8)
# relative lengths of the 24 human chromosomes (1..22, X, Y), Mbp
_CHROM_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80,
              59, 64, 47, 51, 156, 57]


def c5_chromosomes(total=3_100_000_000):
    """Lengths of the 24 chromosomes (each < 2^31), summing to `total`."""
    w = np.array(_CHROM_MBP, dtype=np.float64)
    lens = np.floor(w / w.sum() * total).astype(np.int64)
    lens[0] += total - int(lens.sum())
    return lens


def c5_genome_region(lo, hi, seed=9):
    """Letters [lo, hi) of the genome laid out as one long string (the chromosomes back to back):
    i.i.d. bases, GC 0.41, isolated IUPAC ambiguity letters at rate 1e-3, never within 15 letters
    of each other, no runs of N (assembly gaps are masked out up front)."""
    out = synth.subject_region(lo, hi, 1 << 62, seed=seed, with_n=False)
    B = synth.BLOCK
    for b in range(lo // B, (hi + B - 1) // B):
        g = np.random.Generator(np.random.PCG64([seed, 4242, b]))
        pos = np.flatnonzero(g.random(B, dtype=np.float32) < 1e-3)
        pos = pos[(pos >= 16) & (pos < B - 16)]                # never near a block edge
        if pos.size:
            pos = pos[np.concatenate([[True], np.diff(pos) > 15])]
        let = AMBIGUITY[g.integers(0, AMBIGUITY.size, pos.size)]
        gp = pos + b * B
        keep = (gp >= lo) & (gp < hi)
        out[gp[keep] - lo] = let[keep]
    return out


def c5_dictionary(M=2_000_000, total=3_100_000_000, seed=8, frac_planted=0.10):
    """M patterns of 15-40 letters (concat, widths): i.i.d. random, 10 % cut from 64 evenly spaced
    1 Mi-letter blocks of the genome; where the genome holds an ambiguity letter the pattern gets one
    of the bases it stands for, so those occurrences are found with fixed=FALSE only (in the Trusted
    Band: through the expansion of the ambiguous subject window)."""
    g = np.random.Generator(np.random.PCG64(seed))
    widths = g.integers(C5_MIN_W, C5_MAX_W + 1, M).astype(np.int32)
    offs = np.concatenate([[0], np.cumsum(widths, dtype=np.int64)])
    concat = BASES[g.integers(0, 4, int(offs[-1]), dtype=np.uint8)]
    n_pl = int(M * frac_planted)
    rows = g.choice(M, n_pl, replace=False)
    B = synth.BLOCK
    nb = max(1, total // B - 1)
    blocks = np.arange(0, nb, max(1, -(-nb // 64)), dtype=np.int64)
    blk = blocks[g.integers(0, blocks.size, n_pl)]
    off = g.integers(0, B - C5_MAX_W, n_pl)
    order = np.argsort(blk, kind="stable")
    cur, letters = -1, None
    for k in order.tolist():
        if blk[k] != cur:
            cur = int(blk[k])
            letters = c5_genome_region(cur * B, (cur + 1) * B)
        i = int(rows[k])
        w = letters[off[k]:off[k] + widths[i]].copy()
        for j in np.flatnonzero(~np.isin(w, BASES)).tolist():
            bits = [b for b in (1, 2, 4, 8) if int(w[j]) & b]
            w[j] = bits[int(g.integers(0, len(bits)))]
        concat[offs[i]:offs[i + 1]] = w
    return concat, widths


def c5_pieces(lo, hi, chrom_lens):
    """The stretch [lo, hi) of the back-to-back genome as per-chromosome pieces
    (chromosome index, start inside the chromosome, end inside the chromosome)."""
    out = []
    base = 0
    for c, ln in enumerate(chrom_lens.tolist()):
        a, z = max(lo, base), min(hi, base + ln)
        if z > a:
            out.append((c, a - base, z - base))
        base += ln
    return out
