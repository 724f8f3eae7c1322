"""Python restatement of the reference's matching with indels (TEST INFRASTRUCTURE ONLY: only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this).

The one piece of SURVEY 8f rank 1 the CUDA path still refuses (`with.indels=TRUE` on a
non-preprocessed dictionary: countPDict / whichPDict / vcountPDict / vwhichPDict only,
R/matchPDict.R:171-191) -- restated here, pinned against the reference's own C, and split the way
a device implementation has to split it, so that the next round starts from a checked
specification:

  candidates()   for every subject position j0 independently: the candidate match that
                 _match_pattern_indels (src/match_pattern_indels.c:58-107) would hand to
                 report_provisory_match() -- (start, width, nedit) or nothing.  Per position this
                 is one table lookup (first pattern letter the subject letter matches) and one
                 banded edit distance _nedit_for_Ploffset (src/lowlevel_matching.c:132-213: band
                 2k+1, rows = pattern letters, out-of-subject letters are mismatches, early
                 bailout) -- embarrassingly parallel over (pattern, position).
  best_local()   the sequential "best local matches" filter over the candidates in ascending
                 start order (report_provisory_match, :38-56) -- touches candidates only.

Follows: src/match_pattern_indels.c:24-107, src/lowlevel_matching.c:26-89 (match tables,
_nmismatch_at_Pshift), :95-213 (PROPAGATE_NEDIT, _nedit_for_Ploffset), src/match_pattern.c:85-112
(the indels algorithm is only entered when the pattern is longer than max.mismatch) and
R/match-utils.R:54-57 (min.mismatch must be 0).
"""
import numpy as np


def match_table(fixedP, fixedS):
    """lowlevel_matching.c:26-56 on raw code bytes: returns f(p, s) -> bool."""
    if fixedP and fixedS:
        return lambda p, s: p == s
    if fixedP:
        return lambda p, s: (p & ~s) & 0xFF == 0
    if fixedS:
        return lambda p, s: (~p & s) & 0xFF == 0
    return lambda p, s: (p & s) != 0


def nmismatch_at_pshift(P, S, shift, max_nmis, match):
    """_nmismatch_at_Pshift, lowlevel_matching.c:66-89: letters outside S are mismatches; stops
    as soon as the count exceeds max_nmis."""
    nmis = 0
    for i, p in enumerate(P):
        j = shift + i
        if j < 0 or j >= len(S) or not match(int(p), int(S[j])):
            nmis += 1
            if nmis > max_nmis:
                break
    return nmis


def nedit_for_ploffset(P, S, ploffset, max_nedit, match):
    """_nedit_for_Ploffset, lowlevel_matching.c:132-213 -> (min_nedit, min_width)."""
    if len(P) == 0:
        return 0, 0
    assert max_nedit != 0
    max_nedit_plus1 = max_nedit + 1
    if max_nedit > len(P):
        max_nedit = len(P)
    row_length = 2 * max_nedit + 1
    INF = 1 << 30
    curr = [INF] * row_length
    prev = [INF] * row_length

    def cost(pc, si):
        return int(si < 0 or si >= len(S) or not match(pc, int(S[si])))

    def propagate(curr, B, prev, si, pc):
        nedit = prev[B] + cost(pc, si)
        if B - 1 >= 0 and curr[B - 1] + 1 < nedit:
            nedit = curr[B - 1] + 1
        if B + 1 < row_length and prev[B + 1] + 1 < nedit:
            nedit = prev[B + 1] + 1
        curr[B] = nedit

    min_si = ploffset
    # stage 0
    for b, B in enumerate(range(max_nedit, row_length)):
        curr[B] = b
    # stage 1
    a, pi = 1, 0
    while a < max_nedit:
        pc = int(P[pi])
        prev, curr = curr, prev
        B = max_nedit - a
        curr[B] = a
        B += 1
        si = min_si
        while B < row_length:
            propagate(curr, B, prev, si, pc)
            B += 1
            si += 1
        a += 1
        pi += 1
    # stage 2
    pc = int(P[pi])
    prev, curr = curr, prev
    curr[0] = min_nedit = a
    min_width = 0
    si = min_si
    for B in range(1, row_length):
        propagate(curr, B, prev, si, pc)
        if curr[B] < min_nedit:
            min_nedit = curr[B]
            min_width = si - ploffset + 1
        si += 1
    a += 1
    pi += 1
    # stage 3
    while pi < len(P):
        pc = int(P[pi])
        prev, curr = curr, prev
        min_nedit = a
        min_width = 0
        si = min_si
        for B in range(row_length):
            propagate(curr, B, prev, si, pc)
            if curr[B] < min_nedit:
                min_nedit = curr[B]
                min_width = si - ploffset + 1
            si += 1
        if min_nedit >= max_nedit_plus1:
            break
        pi += 1
        a += 1
        min_si += 1
    return min_nedit, min_width


def candidates(P, S, max_nmis, fixedP=True, fixedS=True):
    """Per subject position, independently: [(start1, width, nedit)] in ascending start order."""
    match = match_table(fixedP, fixedS)
    # byte2offset (src/utils.c:77-97): first pattern position a subject byte matches
    first = {}
    out = []
    for j0 in range(len(S)):
        c0 = int(S[j0])
        if c0 not in first:
            first[c0] = next((n for n, p in enumerate(P) if match(int(p), c0)), None)
        i0 = first[c0]
        if i0 is None:
            continue
        P1 = P[i0 + 1:]
        k = max_nmis - i0
        if k < 0:
            continue
        if k == 0:
            nedit1, width1 = nmismatch_at_pshift(P1, S, j0 + 1, 0, match), len(P1)
        else:
            nedit1, width1 = nedit_for_ploffset(P1, S, j0 + 1, k, match)
        if nedit1 <= k:
            out.append((j0 + 1, width1 + 1, nedit1 + i0))
    return out


def best_local(cands):
    """report_provisory_match + the final flush -> [(start1, width)] as _report_match gets them."""
    out = []
    prov = None                         # (start, end, width, nedit)
    for start, width, nedit in cands:
        end = start + width - 1
        if prov is not None:
            if end > prov[1]:
                out.append((prov[0], prov[2]))
            elif nedit > prov[3]:
                continue
        prov = (start, end, width, nedit)
    if prov is not None:
        out.append((prov[0], prov[2]))
    return out


def match_pattern_indels(P, S, max_nmis, fixedP=True, fixedS=True):
    """_match_pattern_indels: the reported (start, width) pairs."""
    if len(P) == 0:
        raise ValueError("empty pattern")
    return best_local(candidates(np.asarray(P), np.asarray(S), max_nmis, fixedP, fixedS))


def count_indels(patterns, S, max_nmis, fixedP=True, fixedS=True):
    """countPDict(<XStringSet>, S, max.mismatch, with.indels=TRUE): matches per pattern
    (src/match_pattern.c:85-112: a pattern not longer than max.mismatch goes through the naive
    inexact loop instead, and `max_nmis < len(P) - len(S)` reports nothing)."""
    match = match_table(fixedP, fixedS)
    S = np.asarray(S)
    out = []
    for P in patterns:
        P = np.asarray(P)
        if max_nmis < len(P) - len(S):
            out.append(0)
        elif len(P) <= max_nmis:
            lo = 1 - len(P)
            n = sum(1 for sh in range(lo, len(S) - lo - len(P) + 1)
                    if nmismatch_at_pshift(P, S, sh, max_nmis, match) <= max_nmis)
            out.append(n)
        else:
            out.append(len(match_pattern_indels(P, S, max_nmis, fixedP, fixedS)))
    return out
