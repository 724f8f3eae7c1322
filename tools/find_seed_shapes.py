"""Offline search of gapped seed shapes + shift sets for Hamming-distance filtering.

A seed family = one shape (q offsets inside a span) applied at a set T of shifts inside a
window of m letters.  It is complete for (m, k) when for every set of k mismatch positions
at least one t in T has shape+t free of mismatches (gapped q-grams with threshold >= 1,
Burkhardt & Karkkainen).  Indexing every pattern at the shifts of T then finds all
occurrences with <= k mismatches with ONE probe per subject position and |T|*M/4^q
candidates per position.  We minimise |T| (index size, candidates), then maximise q.

Prints rows {m, k, q, span, shape_mask, shifts_mask} for csrc/bsgpu_shapes.inc.

    python tools/find_seed_shapes.py <seconds per (m,k,q)> <k list, e.g. 1,2> [m_lo m_hi]
"""
import itertools
import math
import random
import sys
import time


def greedy_cover(shape_mask, span, m, subsets, best_so_far):
    """Greedy small set of shifts whose clean-sets cover all k-subsets; None if impossible
    or not better than best_so_far."""
    shifts = list(range(m - span + 1))
    full = (1 << m) - 1
    cover = []
    for t in shifts:
        u = full & ~(shape_mask << t)
        bits = 0
        for i, s in enumerate(subsets):
            if (u & s) == s:
                bits |= 1 << i
        cover.append(bits)
    need = (1 << len(subsets)) - 1
    allc = 0
    for c in cover:
        allc |= c
    if allc != need:
        return None
    chosen, covered = [], 0
    while covered != need:
        if len(chosen) + 1 >= best_so_far:
            return None
        t = max(shifts, key=lambda x: bin(cover[x] & ~covered).count("1"))
        chosen.append(t)
        covered |= cover[t]
    return chosen


def search(m, k, q, budget_s, rnd):
    subsets = [sum(1 << x for x in c) for c in itertools.combinations(range(m), k)]
    best = None
    t0 = time.time()
    for span in range(q, m + 1):
        if time.time() - t0 > budget_s:
            break
        n_inner = span - 2
        if n_inner < q - 2:
            continue
        total = math.comb(n_inner, q - 2)
        tries = min(total, 300)
        if total <= tries:
            cands = itertools.combinations(range(1, span - 1), q - 2)
        else:
            cands = (tuple(sorted(rnd.sample(range(1, span - 1), q - 2))) for _ in range(tries))
        for inner in cands:
            mask = 1 | (1 << (span - 1))
            for i in inner:
                mask |= 1 << i
            r = greedy_cover(mask, span, m, subsets, best[0] if best else 99)
            if r is not None and (best is None or len(r) < best[0]):
                best = (len(r), span, mask, sorted(r))
            if time.time() - t0 > budget_s:
                break
    return best


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3]
    m_lo = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    m_hi = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    rnd = random.Random(12345)
    print("// {m, k, q, span, shape_mask, shifts_mask}")
    for k in ks:
        for m in range(m_lo, m_hi + 1):
            best = None
            for q in (12, 11, 10):
                r = search(m, k, q, budget, rnd)
                if r:
                    best = (q,) + r
                    break
            if best:
                q, n, span, mask, shifts = best
                sm = 0
                for t in shifts:
                    sm |= 1 << t
                print(f"{{{m}, {k}, {q}, {span}, 0x{mask:x}ull, 0x{sm:x}ull}}, // {n} shifts {shifts}",
                      flush=True)
