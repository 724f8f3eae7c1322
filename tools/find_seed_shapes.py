"""Offline search of gapped seed shapes for Hamming-distance filtering.

A shape (set of offsets, weight q, span <= m) has threshold >= 1 for (m, k) when for every
set of k mismatch positions inside a window of m letters at least one of the m-span+1
shifts of the shape avoids all of them (Burkhardt & Karkkainen's gapped q-grams).  An index
of the shape at every shift of every pattern then finds all occurrences with <= k
mismatches with ONE probe per subject position.  Prints a C table for bsgpu_qindex.cuh.
"""
import itertools
import random
import sys
import time


def covers(shape_mask, span, m, k):
    full = (1 << m) - 1
    U = [full & ~(shape_mask << t) for t in range(m - span + 1)]
    if k == 1:
        acc = 0
        for u in U:
            acc |= u
        return acc == full
    if k == 2:
        for a in range(m):
            acc = 0
            for u in U:
                if (u >> a) & 1:
                    acc |= u
            if acc != full:
                return False
        return True
    if k == 3:
        for a in range(m):
            Ua = [u for u in U if (u >> a) & 1]
            for b in range(a + 1, m):
                acc = 0
                for u in Ua:
                    if (u >> b) & 1:
                        acc |= u
                if acc != full:
                    return False
        return True
    raise ValueError(k)


def search(m, k, q, budget_s):
    """Smallest-span shape of weight q (most shifts), exhaustive when cheap else random."""
    t0 = time.time()
    for span in range(q, m + 1):
        n_inner = span - 2
        import math
        total = math.comb(n_inner, q - 2) if q >= 2 else 1
        if total <= 300_000:
            it = itertools.combinations(range(1, span - 1), q - 2)
        else:
            rnd = random.Random(1234 + span)
            it = (tuple(sorted(rnd.sample(range(1, span - 1), q - 2))) for _ in range(300_000))
        for inner in it:
            mask = 1 | (1 << (span - 1))
            for i in inner:
                mask |= 1 << i
            if covers(mask, span, m, k):
                return span, mask
            if time.time() - t0 > budget_s:
                return None
    return None


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    print("// {m, k, q, span, mask}")
    for k in (1, 2, 3):
        for m in range(14, 65):
            best = None
            for q in (12, 11, 10):
                r = search(m, k, q, budget)
                if r:
                    best = (q,) + r
                    break
            if best:
                q, span, mask = best
                print(f"{{{m}, {k}, {q}, {span}, 0x{mask:x}ull}},", flush=True)
