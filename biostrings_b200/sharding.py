"""Sharding the PDict path over the GPUs of one box.

A long subject is split by Trusted-Band end position: rank r reports the TB ends n with
lo_r < n <= hi_r and needs the letters [lo_r - (w-1) - max|head|, hi_r + max|tail|] (an
overlap of width(pattern)-1).  "Out of limits" is judged against the whole subject, so
only the first / last rank see the real edges.  Per-pattern counts of the ranks add up
(one all-reduce, the only collective on this path); per-pattern ends concatenated in rank
order stay ascending.  Subject sets are split by sequence: no halo, no overlap.

One process per GPU; torch.distributed supplies the plumbing (NCCL on GPUs, gloo in the
CPU tests of the host-side logic).
"""
import numpy as np


def shard_bounds(n_units, world, rank):
    """Even split of n_units into `world` contiguous parts: (lo, hi) with lo < unit <= hi
    in 1-based units, i.e. 0-based [lo, hi)."""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def halo_of(pdict):
    """(left, right) letters a chunk must carry around its report range for `pdict`."""
    left = right = 0
    for p in pdict.parts():
        hw = int(p.head.widths.max()) if len(p.head) else 0
        tw = int(p.tail.widths.max()) if len(p.tail) else 0
        left = max(left, p.tb_width() - 1 + hw)
        right = max(right, tw)
    return left, right


def chunk_range(lo, hi, subject_len, halo_left, halo_right):
    """Letters [a, b) (0-based) rank needs to report TB ends in (lo, hi]."""
    return max(0, lo - halo_left), min(int(subject_len), hi + halo_right)


def split_set_by_letters(widths, world):
    """Contiguous split of a sequence set balanced by total letters: list of (i0, i1)."""
    widths = np.asarray(widths, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(widths)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world, side="left")))
    cuts.append(len(widths))
    cuts = np.maximum.accumulate(np.minimum(cuts, len(widths))).tolist()
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def all_reduce_counts(counts, group=None):
    """Sum of the per-rank count vectors (torch tensor, in place) -- the one collective."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def sharded_count(count_fn, subject_len, halo, rank, world, fetch, reduce=all_reduce_counts):
    """Host-side driver of a sharded countPDict.

    count_fn(chunk, chunk_start, subject_len, lo, hi) -> per-pattern counts of the TB ends
    in (lo, hi] (torch tensor); fetch(a, b) -> letters [a, b) of the subject.  Returns the
    all-reduced counts.  The device path passes DeviceSubject.from_chunk + match; the CPU
    tests pass the oracle, so that the sharding arithmetic is checked without a GPU.
    """
    lo, hi = shard_bounds(subject_len, world, rank)
    a, b = chunk_range(lo, hi, subject_len, halo[0], halo[1])
    counts = count_fn(fetch(a, b), a, subject_len, lo, hi)
    return reduce(counts)


def sharded_oligo_frequency(count_fn, subject_len, width, rank, world, fetch, reduce=all_reduce_counts):
    """Host-side driver of a sharded oligonucleotideFrequency(x, width, step) of one long sequence
    (SURVEY 8f rank 4 on the 8e sharding): rank r counts the windows whose LAST letter n (1-based)
    lies in (lo_r, hi_r], which needs the width-1 letters in front of lo_r and nothing behind hi_r;
    window starts are judged against `step` in whole-sequence coordinates, so the per-rank tables
    add up to the table of the sequence (one all-reduce of 4^width counters).

    count_fn(chunk, chunk_start, subject_len, lo, hi) -> table of that chunk (torch tensor); the
    device path passes DeviceSubject.from_chunk + bsgpu_oligo_frequency / bsgpu_oligo_count_device,
    the CPU tests pass the oracle.
    """
    lo, hi = shard_bounds(subject_len, world, rank)
    a, b = chunk_range(lo, hi, subject_len, width - 1, 0)
    return reduce(count_fn(fetch(a, b), a, subject_len, lo, hi))
