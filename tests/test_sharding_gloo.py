"""The N>1 path on CPU: two processes (gloo) each count their overlap chunk of one long
subject with the CPU oracle standing in for the device call; the all-reduced counts must
equal the single-process answer, and the ends concatenated in rank order must stay in the
reference's order.  This exercises the host-side sharding arithmetic (halo, report range,
global out-of-limits) that the GPU path uses unchanged."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from biostrings_b200 import sharding
from oracle import rlevel as R
from tests.helpers import random_dna, rng

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_case():
    g = rng(77)
    s = random_dna(g, 6000)
    s[2990:3010] = 15
    pats = [s[i:i + 20].copy() for i in g.integers(0, 5980, 150)]
    pats += [np.concatenate([random_dna(g, 4), s[0:16]]), np.concatenate([s[5985:6000], random_dna(g, 5)])]
    for p in pats[:60]:
        p[int(g.integers(10, 20))] = 1
        p[p == 15] = 2
    for p in pats:
        p[p == 15] = 2
    return s, pats


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    s, pats = _make_case()
    pd = R.PDict(R.SeqSet(pats), tb_start=3, tb_end=10, backend="port")
    tp = pd.threeparts_list[0]
    halo = (10 - 2 - 1 + 2, 10)          # w-1+max|head|, max|tail|

    def count_fn(chunk, chunk_start, subject_len, lo, hi):
        # the oracle has no notion of chunks: pad with a never-matching byte up to the
        # real subject edges so "out of limits" is judged globally, then keep the TB ends
        # of the report range (what bsgpu_subject_chunk does on the device)
        ends = tp.pptb.match_xstring(tp.head_or_none(), tp.tail_or_none(), chunk, 2, 0,
                                     (True, True), R.MATCHES_AS_ENDS)
        counts = np.zeros(len(pats), dtype=np.int64)
        kept = []
        for i, e in enumerate(ends):
            tl = int(tp.tail.widths[i])
            if e is None:
                kept.append([])
                continue
            tb_end = e - tl + chunk_start
            ok = (tb_end > lo) & (tb_end <= hi)
            # a match whose flank left the chunk on a side that is not a subject edge would
            # have been wrong; the halo makes that impossible inside the report range
            counts[i] = int(ok.sum())
            kept.append((e[ok] + chunk_start).tolist())
        out[rank] = kept
        return torch.from_numpy(counts)

    total = sharding.sharded_count(count_fn, len(s), halo, rank, WORLD, lambda a, b: s[a:b])
    if rank == 0:
        out["total"] = total.numpy().tolist()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_count_equals_whole():
    s, pats = _make_case()
    pd = R.PDict(R.SeqSet(pats), tb_start=3, tb_end=10, backend="port")
    want_counts = R.countPDict(pd, s, max_mismatch=2)
    want_ends = R.matchPDict(pd, s, max_mismatch=2).endIndex()
    mgr = tmp.Manager()
    out = mgr.dict()
    tmp.spawn(_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    assert out["total"] == want_counts.tolist()
    for i in range(len(pats)):
        got = out[0][i] + out[1][i]
        assert got == ([] if want_ends[i] is None else want_ends[i].tolist())


def _oligo_worker(rank, port, out, width, step):
    from oracle import letterfreq as LF
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    s, _ = _make_case()

    def count_fn(chunk, chunk_start, subject_len, lo, hi):
        # what a chunk subject makes oligo_freq_kernel do: windows whose last letter (1-based)
        # is in (lo, hi], starts judged against `step` in whole-sequence coordinates
        table = np.zeros(4 ** width, dtype=np.int64)
        for st in range(0, len(chunk) - width + 1):
            g0 = chunk_start + st
            if g0 % step or not (lo < g0 + width <= hi):
                continue
            table += LF.oligo_counts(chunk[st:st + width], width)
        return torch.from_numpy(table)

    total = sharding.sharded_oligo_frequency(count_fn, len(s), width, rank, WORLD, lambda a, b: s[a:b])
    if rank == 0:
        out["total"] = total.numpy().tolist()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("width,step", [(3, 1), (5, 3)])
def test_two_rank_sharded_oligo_frequency_equals_whole(width, step):
    from oracle import letterfreq as LF
    s, _ = _make_case()
    mgr = tmp.Manager()
    out = mgr.dict()
    tmp.spawn(_oligo_worker, args=(_free_port(), out, width, step), nprocs=WORLD, join=True)
    assert out["total"] == LF.oligo_counts(s, width, step).tolist()
