"""Seeded synthetic inputs of the BASELINE.json shapes.

Everything is generated from numpy PCG64 streams keyed by (seed, block) so that any rank
can materialise any region of a long subject without generating the rest: the subject is
a sequence of 1 Mi-letter blocks, block b drawn from PCG64([seed, b]).
"""
import numpy as np

BASES = np.array([1, 2, 4, 8], dtype=np.uint8)
BLOCK = 1 << 20


def _block(seed, b, gc):
    g = np.random.Generator(np.random.PCG64([seed, b]))
    u = g.random(BLOCK, dtype=np.float32)
    at = (1.0 - gc) / 2.0
    # A | C | G | T with P(A)=P(T)=(1-gc)/2, P(C)=P(G)=gc/2
    idx = (u >= at).astype(np.uint8) + (u >= 0.5) + (u >= 0.5 + gc / 2.0)
    return BASES[idx]


def n_runs(chrom_len, chrom_index, seed=3):
    """chr1-shaped runs of N inside one chromosome of `chrom_len` letters: both telomeres,
    one centromere run (~6.8 % of the length) and 37 short gaps.  (start, length) pairs,
    0-based, relative to the chromosome."""
    g = np.random.Generator(np.random.PCG64([seed, 7777, chrom_index]))
    tel = min(10_000, chrom_len // 100)
    runs = [(0, tel), (chrom_len - tel, tel)]
    cen = int(chrom_len * 0.068)
    runs.append((int(chrom_len * 0.484), cen))
    for _ in range(37):
        ln = int(g.integers(max(1, chrom_len // 25_000), max(2, chrom_len // 5_000)))
        runs.append((int(g.integers(tel, chrom_len - tel - ln)), ln))
    return runs


def subject_region(lo, hi, chrom_len, seed=3, gc=0.41, with_n=True, out=None):
    """Letters [lo, hi) (0-based) of the synthetic genome made of back-to-back chromosomes
    of `chrom_len` letters each.  Returns a uint8 array of Biostrings codes."""
    n = hi - lo
    res = out if out is not None else np.empty(n, dtype=np.uint8)
    b0, b1 = lo // BLOCK, (hi + BLOCK - 1) // BLOCK
    for b in range(b0, b1):
        blk = _block(seed, b, gc)
        a, z = max(lo, b * BLOCK), min(hi, (b + 1) * BLOCK)
        res[a - lo:z - lo] = blk[a - b * BLOCK:z - b * BLOCK]
    if with_n:
        c0, c1 = lo // chrom_len, (hi - 1) // chrom_len if hi > lo else lo // chrom_len
        for c in range(c0, c1 + 1):
            base = c * chrom_len
            for st, ln in n_runs(chrom_len, c, seed):
                a, z = max(lo, base + st), min(hi, base + st + ln)
                if z > a:
                    res[a - lo:z - lo] = 15
    return res


PLANT_BLOCKS = 64


def planted_blocks(chrom_len):
    """The 1 Mi-letter blocks of a chromosome the planted patterns are cut from: up to
    PLANT_BLOCKS of them, evenly spaced over the whole chromosome."""
    nb = max(1, chrom_len // BLOCK)
    step = max(1, -(-nb // PLANT_BLOCKS))
    return np.arange(0, nb, step, dtype=np.int64)


def dictionary(M, width, chrom_len, n_chrom, seed=2, frac_planted=0.10, frac_dup=0.01,
               subject_seed=3, gc=0.41, nsub=0, sub_lo=0, with_n=True):
    """M patterns of `width` letters: i.i.d. random, `frac_planted` of them cut from the
    synthetic subject (generated with the same `with_n`) at positions spread over the WHOLE length of every chromosome (uniform
    inside up to 64 evenly spaced 1 Mi-letter blocks per chromosome, so any window of a few Mbp
    and every rank of a sharded run sees hits; windows holding an N stay random), optionally
    with up to `nsub` substitutions at positions >= sub_lo, and `frac_dup` exact duplicates of
    other patterns.  Returns a uint8 array [M, width]."""
    g = np.random.Generator(np.random.PCG64(seed))
    pats = BASES[g.integers(0, 4, size=(M, width), dtype=np.uint8)]
    n_pl = int(M * frac_planted)
    if n_pl:
        rows = g.choice(M, size=n_pl, replace=False)
        chrom = g.integers(0, n_chrom, size=n_pl)
        blocks = planted_blocks(chrom_len)
        blk = blocks[g.integers(0, blocks.size, size=n_pl)]
        span = min(BLOCK, chrom_len) - width
        off = g.integers(0, max(span, 1), size=n_pl)
        order = np.lexsort((blk, chrom))
        cur = (-1, -1)
        letters = None
        for k in order.tolist():
            c, b = int(chrom[k]), int(blk[k])
            if (c, b) != cur:
                a = c * chrom_len + b * BLOCK
                letters = subject_region(a, min(a + BLOCK, (c + 1) * chrom_len), chrom_len, subject_seed, gc, with_n)
                cur = (c, b)
            w = letters[off[k]:off[k] + width]
            if w.size == width and (w != 15).all():
                pats[rows[k]] = w
        if nsub:
            k = g.integers(0, nsub + 1, size=n_pl)
            for r, kk in zip(rows.tolist(), k.tolist()):
                for _ in range(kk):
                    pats[r, int(g.integers(sub_lo, width))] = BASES[int(g.integers(4))]
    n_du = int(M * frac_dup)
    if n_du:
        dst = g.choice(M, size=n_du, replace=False)
        src = g.integers(0, M, size=n_du)
        pats[dst] = pats[src]
    return pats


def reads(n, width, genome_len, seed=7, frac_planted=0.05, probes=None):
    """n reads of `width` letters (i.i.d. random; `frac_planted` of them carry one of
    `probes` at a random offset).  Returns uint8 [n, width]."""
    g = np.random.Generator(np.random.PCG64(seed))
    out = BASES[g.integers(0, 4, size=(n, width), dtype=np.uint8)]
    if probes is not None and frac_planted > 0:
        k = int(n * frac_planted)
        rows = g.choice(n, size=k, replace=False)
        which = g.integers(0, len(probes), size=k)
        off = g.integers(0, width - probes.shape[1] + 1, size=k)
        for r, q, o in zip(rows.tolist(), which.tolist(), off.tolist()):
            out[r, o:o + probes.shape[1]] = probes[q]
    return out
