#!/usr/bin/env python
"""bench.py -- countPDict subject Gbp/s (BASELINE.json metric) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mm 0|2]
                    [--config c2|c3|c4|c5] [--scaling weak|strong] [--collective peer|overlapped|per-step|once]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (config c2, the default; BASELINE.json configs[1]): countPDict of 1M 25-mers (10 %
cut from the subject over its whole length, 1 % duplicates) against a synthetic chr1-shaped
subject (i.i.d. letters, GC 0.41, ~7 % N in 40 runs) of 250 Mbp per GPU (--scaling weak: N x 250
Mbp in total) or 250 Mbp in total (--scaling strong), split by Trusted-Band end with an overlap of
width-1.  Other configurations (BASELINE.json configs[2..4], `biostrings_b200/workloads.py`):
  c3  1M 36-mers, PDict(tb.start=1, tb.end=12), countPDict(max.mismatch=2), 100 Mbp in total
  c4  vcountPDict(collapse=1) of 100k 25-mers against 10M reads of 150 bp, sharded by read
  c5  2M patterns of 15-40 letters, PDict(tb.start=1, tb.width=15), fixed=FALSE, 3.1 Gbp genome of
      24 chromosomes, overlap-sharded

A step = one pass of the hot path over the rank's resident shard of 1-byte letters, treated as a new
batch (derived planes dropped), counting into the rank's device count vector, followed by the
gathering of the counts a countPDict call on N GPUs needs.  --collective peer (default): ONE count
vector in rank 0's memory (bsgpu_shared_counts_*, CUDA IPC), every rank's scan kernels increment it
directly with system-scope reductions over NVLink -- the gather is part of the scan launch, no
collective is launched; overlapped: every step has its own NCCL all-reduce, issued on a side
stream so that it overlaps the scan of the next batch (two count vectors); per-step: the same,
serialised behind the scan on one stream; once: one all-reduce after the last step.  The modes not chosen are reported in `collective_other_modes`.  `value`
times EXACTLY K steps between barrier + synchronize, one CUDA event on each side on the launching
stream, max over ranks.  No L2 flush inside the bracket: every step streams a subject larger than
the 126 MB L2 (config.l2 says so).  `e2e` times the public host-buffer call: pinned host letters
-> sliced H2D overlapped with the scan (bsgpu_count_host) -> all-reduce on the device -> one D2H of
the counts.  `cpu_baseline` runs the reference's own C (oracle/_ref, else our C restatement) on
one host core over a bounded sample; every line carries `parity_with_gpu_on_sample`: that sample
compared bit for bit with the GPU, for the metric's max.mismatch, and for N > 1 also the
all-reduced counts against a recount of every shard on rank 0's GPU alone.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTH = 25


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mm", type=int, default=0, choices=[0, 2],
                    help="max.mismatch of the headline metric (0 = exact, 2 = MTB_PDict); config c2 only")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--collective", default="peer", choices=["peer", "overlapped", "per-step", "once"])
    ap.add_argument("--subject-mbp", type=float, default=250.0, help="config c2: letters per GPU (weak) / in total (strong), Mbp")
    ap.add_argument("--patterns", type=int, default=1_000_000)
    ap.add_argument("--scale", type=float, default=1.0, help="configs c3-c5: shrink the workload (smoke runs)")
    ap.add_argument("--cpu-sample-mbp", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mm2", action="store_true")
    ap.add_argument("--no-recount", action="store_true", help="N > 1: skip the 1-GPU recount of every shard")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.004)
        except Exception as e:                      # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def mark(self):
        """Start of the timed region: only later samples count."""
        deadline = time.time() + 3
        while not self.samples and time.time() < deadline and self.is_alive():
            time.sleep(0.005)
        self.first = len(self.samples)
        self.reasons.clear()

    def summary(self):
        s = sorted(self.samples[getattr(self, "first", 0):]) or sorted(self.samples[-1:])
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "samples": len(s), "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------
# CPU arm: the reference's own C (or our port) on host cores
# ----------------------------------------------------------------------------------------

def cpu_backend():
    from oracle import ref
    if ref.available():
        ref.lib()               # mapped in this process (and inherited by forked workers)
        return "reference", ref.RefPPTB, ref
    from oracle import port
    return "port", port.PortPPTB, port


def cpu_count_exact(PPTB, mod, pats, sample):
    """Build the whole-pattern ACtree2 index the way PDict() does and count one sample.
    Returns (counts, build_seconds, match_seconds)."""
    M = pats.shape[0]
    t0 = time.perf_counter()
    pptb = PPTB("ACtree2", pats.reshape(-1), np.full(M, WIDTH, np.int32), None)
    h2l = pptb.high2low.copy()
    pptb.blank_dups()
    t1 = time.perf_counter()
    c = pptb.match_xstring(None, None, sample, 0, 0, (True, True), mod.MATCHES_AS_COUNTS)
    t2 = time.perf_counter()
    dup = h2l != -2147483648
    c = c.copy()
    c[dup] = c[h2l[dup] - 1]
    pptb.close()
    return c, t1 - t0, t2 - t1


_WORKER = {}


def _cpu_worker_init(pats, chrom_len):
    """One per host core: builds the reference's ACtree2 index of the dictionary once (the tree
    then warms up over the steps exactly as it does over one long subject)."""
    _, PPTB, mod = cpu_backend()
    M = pats.shape[0]
    pptb = PPTB("ACtree2", pats.reshape(-1), np.full(M, WIDTH, np.int32), None)
    pptb.blank_dups()
    _WORKER.update(pptb=pptb, mod=mod, chrom_len=chrom_len)


def _cpu_worker(job):
    lo, hi = job
    from biostrings_b200 import synth
    sample = synth.subject_region(lo, hi, _WORKER["chrom_len"])
    t0 = time.perf_counter()
    c = _WORKER["pptb"].match_xstring(None, None, sample, 0, 0, (True, True),
                                      _WORKER["mod"].MATCHES_AS_COUNTS)
    return time.perf_counter() - t0, int(c.sum())


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the same workload on all
    host cores (the reference is single-threaded: one process per core, each on its own
    slice of the subject; the index is built once per process, outside the timed part).
    A step is a bounded sample: the slice per core is sized so that warmup + steps together
    spend about a minute matching (the whole run ends within a few minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mpc
    from biostrings_b200 import synth
    kind, _, _ = cpu_backend()          # also maps oracle/_ref/libbiostrings_ref.so in this (parent) process
    cores = max(1, min(os.cpu_count() or 1, 32))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    chrom_len = int(args.subject_mbp * 1e6)
    pats = synth.dictionary(args.patterns, WIDTH, chrom_len, 1)
    nsteps = args.warmup + args.steps
    # 1-4 Mbp/s per core: about a minute of matching in total
    per_core = int(min(1e6, max(2e4, 40 * 4e6 / max(nsteps, 1))))
    span = max(1, (chrom_len - 40_000) // (cores * per_core))    # distinct step windows available
    times = []
    ctx = mpc.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(pats, chrom_len)) as pool:
        for step in range(nsteps):
            base = 20_000 + (step % span) * cores * per_core
            jobs = [(base + i * per_core, base + (i + 1) * per_core + WIDTH - 1) for i in range(cores)]
            res = pool.map(_cpu_worker, jobs, chunksize=1)
            if step >= args.warmup:
                times.append(max(r[0] for r in res))        # slowest core's matching time
    ms = 1e3 * sum(times) / len(times)
    value = cores * per_core / (ms * 1e-3) / 1e9
    line = {"metric": "countPDict subject Gbp/s (1M 25-mers, max.mismatch 0)", "value": value,
            "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "impl": "reference",
            "config": c2_config(args, world),
            "cpu_baseline": {"value": value, "unit": "Gbp/s", "cores": cores, "kind": kind,
                             "sample": f"{cores} x {per_core / 1e6:g} Mbp slices of the subject per step, "
                                       "one single-threaded process per core, ACtree2 index built once per "
                                       "process outside the timed part"},
            "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def c2_config(args, world):
    per = args.subject_mbp if args.scaling == "weak" else args.subject_mbp / world
    return {"workload": f"countPDict, {args.patterns} x {WIDTH}-mers (Trusted Band = whole pattern, "
                        f"max.mismatch={args.mm}), synthetic chr1-shaped subject of "
                        f"{per:g} Mbp per GPU ({per * world:g} Mbp total)",
            "patterns": args.patterns, "pattern_width": WIDTH, "max_mismatch": args.mm,
            "subject_letters_per_gpu": int(per * 1e6),
            "sharding": f"{world} overlap chunk(s) by Trusted-Band end, halo = width-1",
            "collective": COLLECTIVE_TEXT[COLLECTIVE_USED.get("mode") or args.collective] if world > 1 else "none",
            "l2": "inputs larger than L2: every step streams the subject chunk through the 126 MB L2 "
                  "(no flush inside the bracket)" if per * 1e6 > 126e6 else
                  "the rank's chunk is smaller than the 126 MB L2: consecutive steps may find it there (strong scaling)",
            "step": "new batch of 1-byte letters resident in HBM -> scan into the count vector (max.mismatch 0: the "
                    "extended-seed scan reads the letters directly and confirms its candidates; otherwise pack to 2-bit "
                    "planes first); --collective peer (default): the vector is zeroed once before the K steps and must "
                    "hold K x the counts of one step afterwards (checked), the other modes zero it every step"}


COLLECTIVE_USED = {}        # "mode": what timed_steps fell back to when the peer mapping is not available

COLLECTIVE_TEXT = {
    "peer": "none launched: one int32 count vector in rank 0's memory (CUDA IPC, bsgpu_shared_counts_*), every rank's scan "
            "kernels add their matches to it with system-scope reductions over NVLink peer memory, i.e. the gather of the "
            "counts is fused into the scan launch; zeroed once before the K steps (K x the per-step counts checked)",
    "overlapped": "all-reduce of the int32 counts of EVERY step (what one countPDict call on N GPUs needs), on a side "
                  "stream: it overlaps the scan of the next batch (two count vectors)",
    "per-step": "all-reduce of the int32 counts after every step, serialised behind the scan on the launching stream",
    "once": "one all-reduce after the last step"}


# ----------------------------------------------------------------------------------------
# Workloads: what a rank holds, one device step, one end-to-end step, the parity check
# ----------------------------------------------------------------------------------------

class Env:
    """torch / distributed / library plumbing of one rank."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from biostrings_b200 import _lib
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torchrun")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the PDict path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.numa = bind_to_gpu_numa_node(self.local)
        self.L = _lib.lib()
        _lib.check(self.L.bsgpu_set_device(self.local))
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = torch.device("cuda", self.local)
        self.stream = torch.cuda.current_stream()
        self.sptr = C.c_void_p(self.stream.cuda_stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def shared_counts(self, M):
        """One int32[M] count vector for all ranks, in rank 0's memory: returns (device pointer valid
        in this process, torch view of it on rank 0 / None elsewhere).  Cached per M."""
        from biostrings_b200 import _lib
        torch = self.torch
        if not hasattr(self, "_shared"):
            self._shared = {}
        if M in self._shared:
            return self._shared[M]
        ptr = C.c_void_p()
        h = torch.zeros(64, dtype=torch.uint8, device=self.dev)
        ok = 1
        if self.rank == 0:
            buf = (C.c_ubyte * 64)()
            if self.L.bsgpu_shared_counts_create(M, C.byref(ptr), buf) == 0:
                h.copy_(torch.tensor(list(buf), dtype=torch.uint8))
            else:
                ok = 0
        if self.world > 1:
            self.dist.broadcast(h, 0)
        view = None
        if self.rank != 0:
            buf = (C.c_ubyte * 64)(*h.cpu().tolist())
            if self.L.bsgpu_shared_counts_open(buf, C.byref(ptr)) != 0:
                ok = 0
        elif ok:
            view = torch.as_tensor(_DevArray(ptr.value, M), device=self.dev)
        # every rank must have the vector (CUDA IPC + peer access), or nobody uses it
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self._shared[M] = None
            return None
        self._shared[M] = (ptr.value, view)
        return self._shared[M]

    def broadcast0(self, t):
        if self.world > 1:
            self.dist.broadcast(t, 0)
        return t

    def all_reduce(self, t, op=None):
        if self.world > 1:
            self.dist.all_reduce(t, op=op or self.dist.ReduceOp.SUM)
        return t

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.all_reduce(t, self.dist.ReduceOp.MAX if self.world > 1 else None)
        return t.item()


class _DevArray:
    """A raw device pointer as something torch.as_tensor understands (rank 0's own allocation)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def bind_to_gpu_numa_node(local):
    """Best effort: run this rank on the CPUs of its GPU's NUMA node BEFORE it allocates its pinned
    host buffer (the e2e leg is one host-memory -> PCIe stream per GPU).  BENCH_NUMA=0 disables it."""
    if os.environ.get("BENCH_NUMA", "1") == "0":
        return None
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            for part in f.read().strip().split(","):
                if part:
                    lo_, _, hi_ = part.partition("-")
                    cpus.update(range(int(lo_), int(hi_ or lo_) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if len(allowed) < 2:
            return None
        os.sched_setaffinity(0, allowed)
        return {"gpu_pci": bus, "numa_node": node, "cpus": len(allowed)}
    except Exception:           # no sysfs, no such attribute, not permitted: leave the rank where it is
        return None


class Pieces:
    """A rank's share of one or more long subjects as overlap chunks: pinned host letters, the
    resident device chunks, and the two ways of counting them."""

    def __init__(self, env, parts, mm, fixed):
        from biostrings_b200 import matchpdict as mp
        self.env, self.mp, self.parts, self.mm, self.fixed = env, mp, parts, mm, fixed
        self.parr = mp._parts_array(parts)
        self.items = []                 # (pinned tensor, numpy view, chunk_start, subject_len, lo, hi, DeviceSubject)
        self.letters = 0
        self.h2d_bytes = 0

    def add(self, fill, a, b, subject_len, lo, hi):
        torch = self.env.torch
        host = torch.empty(b - a, dtype=torch.uint8, pin_memory=True)
        chunk = host.numpy()
        fill(chunk)
        ds = self.mp.DeviceSubject.from_chunk(chunk, a, subject_len, lo, hi)
        self.items.append((host, chunk, a, subject_len, lo, hi, ds))
        self.letters += hi - lo
        self.h2d_bytes += b - a

    def device_step(self, counts):
        from biostrings_b200 import _lib
        L, nl = self.env.L, C.c_int(0)
        for it in self.items:
            _lib.check(L.bsgpu_subject_repack(it[6].handle, self.env.sptr))
            _lib.check(L.bsgpu_count_device(self.parr, len(self.parts), it[6].handle, self.mm, 0,
                                            int(self.fixed[0]), int(self.fixed[1]),
                                            C.c_void_p(counts if isinstance(counts, int) else counts.data_ptr()),
                                            self.env.sptr, C.byref(nl)))

    def e2e_step(self, counts):
        for it in self.items:
            self.mp.count_parts_host(self.parts, it[1], self.mm, 0, self.fixed, chunk_start=it[2],
                                     subject_len=it[3], report=(it[4], it[5]),
                                     d_counts=counts if isinstance(counts, int) else counts.data_ptr(),
                                     stream=self.env.stream.cuda_stream)


def timed_steps(env, pieces, counts, nsteps, collective):
    """EXACTLY nsteps steps bracketed by barrier + synchronize and one CUDA event on each side of
    the launching stream; max over ranks.  collective: "overlapped" | "per-step" | "once" (see the
    module docstring).  Returns (total_ms, {kernel: mean ms per launch}, launches, counts of ONE step
    summed over ranks as a numpy array)."""
    from biostrings_b200 import _lib
    torch, L = env.torch, env.L
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    one = torch.zeros_like(counts)
    if env.world == 1 and collective == "overlapped":
        collective = "per-step"
    if collective == "peer":
        # every rank counts straight into rank 0's vector: nothing to launch but the scans.  The
        # vector is zeroed once, before the bracket, and accumulates the K steps (every step does
        # its whole gather -- each match is one reduction over NVLink -- only the 4 MB memset per
        # call is left out, which would need a barrier of its own between the steps)
        M = counts.numel()
        shared = (counts.data_ptr(), counts) if env.world == 1 else env.shared_counts(M)    # N = 1: the same protocol, local vector
        if shared is None:              # no peer mapping on this box: the NCCL path
            COLLECTIVE_USED["mode"] = "overlapped"
            return timed_steps(env, pieces, counts, nsteps, "overlapped")
        ptr, view = shared
        if view is not None:
            view.zero_()
        L.bsgpu_kernel_timing(1)
        l0 = L.bsgpu_launch_count()
        env.barrier()
        ev0.record(env.stream)
        for _ in range(nsteps):
            pieces.device_step(ptr)
        ev1.record(env.stream)
        env.barrier()
        launches = int(L.bsgpu_launch_count() - l0)
        acc = {}
        for name, ms in _lib.kernel_times():
            acc.setdefault(name, []).append(ms)
        L.bsgpu_kernel_timing(0)
        total_ms = env.max_over_ranks(ev0.elapsed_time(ev1))
        if view is not None:
            one.copy_(view)
        env.broadcast0(one)
        tot = one.cpu().numpy().astype(np.int64)
        if env.world == 1:
            counts.zero_()
        assert (tot % nsteps == 0).all(), "every step must find the same matches"
        return total_ms, {k: sum(v) / len(v) for k, v in acc.items()}, launches, tot // nsteps
    bufs = [counts, torch.zeros_like(counts)]
    comm = torch.cuda.Stream() if collective == "overlapped" else None
    scanned = [torch.cuda.Event() for _ in range(2)]
    reduced = [torch.cuda.Event() for _ in range(2)]
    L.bsgpu_kernel_timing(1)
    l0 = L.bsgpu_launch_count()
    env.barrier()
    ev0.record(env.stream)
    if collective == "overlapped":
        for k in range(nsteps):
            buf = bufs[k & 1]
            if k >= 2:
                env.stream.wait_event(reduced[k & 1])       # the all-reduce of step k - 2 is done with buf
            buf.zero_()
            pieces.device_step(buf)
            scanned[k & 1].record(env.stream)
            with torch.cuda.stream(comm):
                comm.wait_event(scanned[k & 1])
                env.all_reduce(buf)
                reduced[k & 1].record(comm)
        env.stream.wait_event(reduced[0])
        env.stream.wait_event(reduced[1])
        one.copy_(bufs[(nsteps - 1) & 1])
    elif collective == "per-step":
        for _ in range(nsteps):
            counts.zero_()
            pieces.device_step(counts)
            env.all_reduce(counts)
        one.copy_(counts)
    else:
        counts.zero_()
        for _ in range(nsteps):
            pieces.device_step(counts)
        env.all_reduce(counts)
    ev1.record(env.stream)
    env.barrier()
    launches = int(L.bsgpu_launch_count() - l0)
    acc = {}
    for name, ms in _lib.kernel_times():
        acc.setdefault(name, []).append(ms)
    L.bsgpu_kernel_timing(0)
    total_ms = env.max_over_ranks(ev0.elapsed_time(ev1))
    if collective != "once":
        per_step = one.cpu().numpy().astype(np.int64)
    else:
        tot = counts.cpu().numpy().astype(np.int64)
        assert (tot % nsteps == 0).all(), "every step must find the same matches"
        per_step = tot // nsteps
    return total_ms, {k: sum(v) / len(v) for k, v in acc.items()}, launches, per_step


def e2e_leg(env, pieces, counts, M, n, total_letters, collective="overlapped"):
    """The public host-buffer path, n timed calls after one warm-up: pinned letters -> sliced H2D
    overlapped with the scan -> the counts of all ranks in one vector -> ONE D2H of the counts into
    pinned memory.  collective "peer" (N > 1): every rank's scans add into rank 0's vector over
    NVLink, a barrier says that all of them are done, rank 0 copies the vector out; otherwise an
    NCCL all-reduce on the device before the copy.  Also times the bare pinned H2D copy of the same
    bytes on all ranks at once (the ceiling this box gives the path)."""
    torch = env.torch
    out = torch.empty(M, dtype=torch.int32, pin_memory=True)
    peer = collective == "peer" and env.world > 1 and env.shared_counts(M) is not None
    times = []
    if peer:
        ptr, view = env.shared_counts(M)
    for k in range(n + 1):
        if peer and view is not None:
            view.zero_()                # (before the bracket: zeroing inside it would need a barrier of its own)
        env.barrier()
        t0 = time.perf_counter()
        if peer:
            pieces.e2e_step(ptr)
            env.barrier()
            if view is not None:
                out.copy_(view, non_blocking=True)
        else:
            counts.zero_()
            pieces.e2e_step(counts)
            env.all_reduce(counts)
            out.copy_(counts, non_blocking=True)
        torch.cuda.synchronize()
        if k > 0:
            times.append(time.perf_counter() - t0)
    if peer:                            # the other ranks get the vector too (they compare it with their steps)
        if view is not None:
            counts.copy_(view)
        env.broadcast0(counts)
        out.copy_(counts)
        torch.cuda.synchronize()
    te = env.max_over_ranks(sum(times) / len(times))
    # the H2D ceiling: the same pinned buffers, one plain copy each, all ranks together
    dst = [torch.empty_like(it[0], device=env.dev) for it in pieces.items]
    ct = []
    for k in range(4):
        env.barrier()
        t0 = time.perf_counter()
        for d, it in zip(dst, pieces.items):
            d.copy_(it[0], non_blocking=True)
        torch.cuda.synchronize()
        if k > 0:
            ct.append(time.perf_counter() - t0)
    tc = env.max_over_ranks(sum(ct) / len(ct))
    del dst
    return {"value": total_letters / te / 1e9, "unit": "Gbp/s",
            "h2d_bytes_per_step": int(pieces.h2d_bytes), "d2h_bytes_per_step": int(4 * M),
            "ms_per_step": 1e3 * te,
            "h2d_ceiling": {"ms": 1e3 * tc, "GB_s_per_gpu": pieces.h2d_bytes / tc / 1e9,
                            "note": "bare cudaMemcpyAsync of the same pinned bytes, all ranks at once, max over ranks",
                            "e2e_over_ceiling": te / tc},
            "call": "bsgpu_count_host (countPDict on host letters: 16 MB slices copied on a copy stream, slice k "
                    "scanned while slice k+1 is in flight, persistent staging buffer) -> " +
                    ("counts added to rank 0's vector over NVLink by the scans themselves, barrier" if peer else
                     "all-reduce on the device") + " -> one D2H of the int32 counts into pinned memory",
            "host_binding": env.numa if env.numa is not None else "none (topology not readable or BENCH_NUMA=0)"}, \
        out.numpy().astype(np.int64)


def recount_on_rank0(env, shards, parts, mm, fixed, M):
    """All shards of the run counted one after the other on rank 0's GPU alone (host letters ->
    bsgpu_count_host): what the all-reduced counts of the N ranks must equal."""
    from biostrings_b200 import matchpdict as mp
    torch = env.torch
    acc = torch.zeros(M, dtype=torch.int32, device=env.dev)
    for fill, a, b, subject_len, lo, hi in shards:
        chunk = np.empty(b - a, dtype=np.uint8)
        fill(chunk)
        mp.count_parts_host(parts, chunk, mm, 0, fixed, chunk_start=a, subject_len=subject_len,
                            report=(lo, hi), d_counts=acc.data_ptr(), stream=env.stream.cuda_stream)
        torch.cuda.synchronize()
    return acc.cpu().numpy().astype(np.int64)


# ----------------------------------------------------------------------------------------
# config c2: the headline
# ----------------------------------------------------------------------------------------

def run_c2(args, env):
    import biostrings_b200 as bs
    from biostrings_b200 import sharding, synth
    torch, L, rank, world = env.torch, env.L, env.rank, env.world
    M = args.patterns
    weak = args.scaling == "weak"
    chrom_len = int(args.subject_mbp * 1e6)
    subject_len = chrom_len * world if weak else chrom_len
    pats = synth.dictionary(M, WIDTH, chrom_len, world if weak else 1)
    dset = bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(M, WIDTH, np.int32))

    def build(mm):
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pd = bs.PDict(dset, max_mismatch=None if mm == 0 else mm)
        return pd, time.perf_counter() - t0

    def shard(r):
        lo, hi = sharding.shard_bounds(subject_len, world, r)
        a, b = sharding.chunk_range(lo, hi, subject_len, WIDTH - 1, 0)
        return (lambda out, a=a, b=b: synth.subject_region(a, b, chrom_len, out=out)), a, b, subject_len, lo, hi

    pdict, t_pdict = build(args.mm)
    parts = pdict.parts()
    pieces = Pieces(env, parts, args.mm, (True, True))
    pieces.add(*shard(rank))
    counts = torch.zeros(M, dtype=torch.int32, device=env.dev)
    per_step = args.collective
    total_letters = subject_len

    def step():
        counts.zero_()
        pieces.device_step(counts)
        env.all_reduce(counts)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()                                  # builds the engine's device index (cached in the PDict3Parts)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    sampler = ClockSampler(env.local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    env.barrier()
    sampler.mark()
    total_ms, kms, launches, hits = timed_steps(env, pieces, counts, args.steps, per_step)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    plan = L.bsgpu_last_plan().decode()
    ms_per_step = total_ms / args.steps
    value = total_letters / (ms_per_step * 1e-3) / 1e9
    other = None
    if world > 1:
        other = []
        n_o = max(3, min(args.steps, 50))
        for mode in ("peer", "overlapped", "per-step", "once"):
            if mode != args.collective:
                o_ms, _, _, o_hits = timed_steps(env, pieces, counts, n_o, mode)
                other.append({"collective": mode, "steps": n_o, "ms_per_step": o_ms / n_o,
                              "value": total_letters / (o_ms / n_o * 1e-3) / 1e9,
                              "same_counts": bool((o_hits == hits).all())})

    # ---- the other half of the metric: max.mismatch = 2 on the same dictionary (MTB_PDict) ----
    mm2 = None
    if args.mm == 0 and not args.no_mm2:
        pd2, _ = build(2)
        p2 = Pieces(env, pd2.parts(), 2, (True, True))
        p2.items = pieces.items             # the same resident chunk
        for _ in range(2):
            counts.zero_()
            p2.device_step(counts)
        env.barrier()
        n2 = max(3, min(args.steps, 10))
        tot2, kms2, l2, hits2 = timed_steps(env, p2, counts, n2, per_step)
        mm2 = {"metric": "countPDict subject Gbp/s (1M 25-mers, max.mismatch 2)",
               "value": total_letters / (tot2 / n2 * 1e-3) / 1e9, "unit": "Gbp/s",
               "steps": n2, "ms_per_step": tot2 / n2, "kernel_ms": kms2,
               "gpu_launches": l2, "matches_total": int(hits2.sum()),
               "plan": L.bsgpu_last_plan().decode()}
        if rank == 0 and not args.no_cpu_baseline:
            mm2["parity_with_gpu_on_sample"] = sample_parity(args, bs, pd2, pats, chrom_len, 2, 2e6)
        if world > 1 and not args.no_recount:
            if rank == 0:
                want = recount_on_rank0(env, [shard(r) for r in range(world)], pd2.parts(), 2, (True, True), M)
                mm2["all_reduced_equals_1gpu_recount"] = bool((want == hits2).all())
            env.barrier()
        del pd2, p2

    # ---- matchPDict (MATCHES_AS_ENDS) on the resident chunk: hits as a CSR on the host ------
    ends_leg = None
    if args.mm == 0 and not args.no_e2e:
        from biostrings_b200 import _lib, matchpdict as mp
        times = []
        for k in range(6):
            env.barrier()
            t0 = time.perf_counter()
            off_e, ends_e = mp.match_parts(parts, pieces.items[0][6], 0, 0, (True, True), _lib.MATCHES_AS_ENDS)
            torch.cuda.synchronize()
            if k > 0:
                times.append(time.perf_counter() - t0)
            nhits = int(ends_e.size)
            del off_e, ends_e               # the result's pinned blocks go back to the library before the next call
        te = sum(times) / len(times)
        ends_leg = {"call": "matchPDict on the resident chunk (ends as a CSR on the host)",
                    "ms": 1e3 * te, "hits_per_rank": nhits, "Mhits_per_s": nhits / te / 1e6,
                    "Gbp_s_per_gpu": pieces.letters / te / 1e9}

    e2e = None
    if not args.no_e2e:
        e2e, e2e_counts = e2e_leg(env, pieces, counts, M, max(2, min(args.steps, 5)), total_letters, args.collective)
        e2e["counts_equal_device_steps"] = bool((e2e_counts == hits).all())

    recount_ok = None
    if world > 1 and not args.no_recount:
        if rank == 0:
            want = recount_on_rank0(env, [shard(r) for r in range(world)], parts, args.mm, (True, True), M)
            recount_ok = bool((want == hits).all())
        env.barrier()
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel, peaks measured in this run --------------------------
    peak, peak_src = peaks()
    dominant = max(kms, key=kms.get)
    dom_ms = kms[dominant]
    alg_bytes = float(pieces.letters)              # 1.0 B per subject letter
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            ent = json.load(f).get(dominant, {})
        traffic = ent.get("dram_bytes_per_launch")
        traffic_note = {k: ent[k] for k in ("capture", "note", "dram_bytes_per_launch_cold") if k in ent} or None
    cg, gg = C.c_double(), C.c_double()
    in_run = None
    if L.bsgpu_probe_peaks(C.byref(cg), C.byref(gg), 32 << 20) == 0:
        in_run = {"copy_GB_s": cg.value, "l2_gather_Gsectors_s_32MB": gg.value,
                  "how": "bsgpu_probe_peaks in this process after the timed region: 1 GiB copy kernel (read + "
                         "write bytes), random 32-byte sectors from a 32 MiB table, best of a few launch shapes"}
    stride_s = 1
    for tok in plan.split(" "):
        if tok.startswith("S="):
            stride_s = int(tok[2:])
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_capture": traffic_note, "kernel": dominant,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
            "kernel_ms": kms,
            "kernel_ms_source": "CUDA events recorded by the library around each of its launches, on the "
                                "launching stream, inside the timed region (bsgpu_kernel_timing); mean per launch",
            "step_frac_of_hbm_peak": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
            "peaks_measured_in_this_run": in_run,
            "plan": plan,
            "note": "algorithmic bytes = 1.0 B per subject letter (the reference's 1-byte encoding, SURVEY 8d); "
                    "traffic = dram__bytes_read+write of that kernel from the committed ncu capture "
                    "(profiles/traffic.json, regenerated by tools/summarize_ncu.py)"}
    if in_run and args.mm == 0:
        rate = pieces.letters / stride_s / (dom_ms * 1e-3) / 1e9
        roof["l2_gather"] = {"achieved_Gprobes_s": rate, "peak": in_run["l2_gather_Gsectors_s_32MB"],
                             "frac": rate / in_run["l2_gather_Gsectors_s_32MB"], "probes_per_letter": 1.0 / stride_s,
                             "peak_source": "measured in this run (peaks_measured_in_this_run)"}

    cpu, parity = None, None
    if not args.no_cpu_baseline:
        parity = sample_parity(args, bs, pdict, pats, chrom_len, args.mm,
                               args.cpu_sample_mbp * 1e6 if args.mm == 0 else 2e6)
        cpu = parity.pop("cpu_baseline")
    line = {"metric": f"countPDict subject Gbp/s (1M 25-mers, max.mismatch {args.mm})",
            "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": c2_config(args, world), "clocks": sampler.summary(),
            "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "parity_with_gpu_on_sample": parity,
            "all_reduced_equals_1gpu_recount": recount_ok,
            "collective_other_modes": other,
            "mm2": mm2, "matches_total": int(hits.sum()), "match_ends": ends_leg,
            "index_build": {"pdict_s": t_pdict, "first_call_s": t_first,
                            "note": "PDict() = host preprocessing (Trusted-Band hash, duplicates) + upload; the "
                                    "first call also builds and uploads the engine's own index (filter and entry "
                                    "table), cached in the PDict3Parts objects"}}
    return line


def sample_parity(args, bs, pdict, pats, chrom_len, mm, nletters):
    """The reference's own C on one host core over a window of the subject, compared bit for bit
    with countPDict of the same window on the GPU.  mm = 0: the whole-pattern ACtree2 (timed: the
    CPU baseline); mm = 2: the three Trusted Bands of PDict(max.mismatch=2) through oracle/rlevel."""
    from biostrings_b200 import synth
    kind, PPTB, mod = cpu_backend()
    n = int(min(nletters, chrom_len - 20_000))
    s_lo = 20_000
    sample = synth.subject_region(s_lo, s_lo + n, chrom_len)
    M = pats.shape[0]
    if mm == 0:
        want, t_build, t_match = cpu_count_exact(PPTB, mod, pats, sample)
        got = bs.countPDict(pdict, sample)
        note = (f"first {n / 1e6:g} Mbp after the telomere of the same subject, same {M}-pattern dictionary; "
                f"ACtree2 index build {t_build:.2f} s not included, cold tree; {t_match:.2f} s of matching")
    else:
        from oracle import rlevel as R
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            od = R.PDict(R.SeqSet(concat=pats.reshape(-1), widths=np.full(M, WIDTH, np.int32)),
                         max_mismatch=mm, backend="ref" if kind == "reference" else "port")
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        want = R.countPDict(od, sample, max_mismatch=mm)
        t_match = time.perf_counter() - t0
        got = bs.countPDict(pdict, sample, max_mismatch=mm)
        note = (f"first {n / 1e6:g} Mbp after the telomere, same dictionary as PDict(max.mismatch={mm}): the "
                f"reference's {mm + 1} Trusted Bands + union (oracle/rlevel over oracle/_ref); {t_match:.2f} s of matching")
    ok = bool((np.asarray(got) == np.asarray(want)).all())
    return {"ok": ok, "max_mismatch": mm, "sample_letters": n, "sample_matches": int(np.asarray(want).sum()),
            "cpu_baseline": {"value": n / t_match / 1e9, "unit": "Gbp/s", "cores": 1, "kind": kind, "sample": note,
                             "parity_with_gpu_on_sample": ok, "sample_matches": int(np.asarray(want).sum())}}


# ----------------------------------------------------------------------------------------
# configs c3 / c5: other dictionaries over overlap chunks; c4: reads sharded by read
# ----------------------------------------------------------------------------------------

def run_c3_c5(args, env):
    import biostrings_b200 as bs
    from biostrings_b200 import sharding, workloads as W
    from oracle import ref, rlevel as R
    torch, L, rank, world = env.torch, env.L, env.rank, env.world
    sc = args.scale
    if args.config == "c3":
        n = int(100e6 * sc)
        M = max(1000, int(1e6 * sc))
        pats = W.c3_dictionary(M, n)
        dkw = dict(concat=pats.reshape(-1), widths=np.full(M, W.C3_WIDTH, np.int32))
        pkw, mm, fixed, maxw = dict(tb_start=1, tb_end=W.C3_TB_END), 2, (True, True), W.C3_WIDTH
        chroms = np.array([n], dtype=np.int64)
        region = lambda lo, hi: W.c3_subject(lo, hi, n)                           # noqa: E731
        label = (f"countPDict(max.mismatch=2), {M} x 36-mers, PDict(tb.start=1, tb.end=12), "
                 f"{n / 1e6:g} Mbp subject in total (SolexaYi2 shape)")
        metric = "countPDict subject Gbp/s (1M 36-mers, tb.end=12, max.mismatch 2)"
    else:
        total = int(3.1e9 * sc)
        M = max(1000, int(2e6 * sc))
        concat, widths = W.c5_dictionary(M, total)
        dkw = dict(concat=concat, widths=widths)
        pkw, mm, fixed, maxw = dict(tb_start=1, tb_width=W.C5_TB_WIDTH), 0, (False, False), W.C5_MAX_W
        chroms = W.c5_chromosomes(total)
        region = W.c5_genome_region
        label = (f"countPDict(fixed=FALSE), {M} patterns of 15-40 letters, PDict(tb.start=1, tb.width=15), "
                 f"{total / 1e9:g} Gbp genome of 24 chromosomes with IUPAC letters at rate 1e-3")
        metric = "countPDict subject Gbp/s (2M 15-40-mers, tb.width=15, fixed=FALSE)"
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pdict = bs.PDict(bs.DNAStringSet(**dkw), **pkw)
    t_pdict = time.perf_counter() - t0
    parts = pdict.parts()
    halo = sharding.halo_of(pdict)
    genome = int(chroms.sum())
    starts = np.concatenate([[0], np.cumsum(chroms)])

    def shards_of(r):
        g_lo, g_hi = sharding.shard_bounds(genome, world, r)
        out = []
        for c, lo, hi in W.c5_pieces(g_lo, g_hi, chroms):
            a, b = sharding.chunk_range(lo, hi, int(chroms[c]), halo[0], halo[1])
            base = int(starts[c])
            out.append((lambda o, a=a, b=b, base=base: o.__setitem__(slice(None), region(base + a, base + b)),
                        a, b, int(chroms[c]), lo, hi))
        return out

    pieces = Pieces(env, parts, mm, fixed)
    for sh in shards_of(rank):
        pieces.add(*sh)
    counts = torch.zeros(M, dtype=torch.int32, device=env.dev)

    def step():
        counts.zero_()
        pieces.device_step(counts)
        env.all_reduce(counts)

    step()
    sampler = ClockSampler(env.local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    env.barrier()
    sampler.mark()
    steps = min(args.steps, 50)
    total_ms, kms, launches, hits = timed_steps(env, pieces, counts, steps, args.collective)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    plan = L.bsgpu_last_plan().decode()
    ms_per_step = total_ms / steps
    e2e = None
    if not args.no_e2e:
        e2e, e2e_counts = e2e_leg(env, pieces, counts, M, 3, genome, args.collective)
        e2e["counts_equal_device_steps"] = bool((e2e_counts == hits).all())
    recount_ok = None
    if world > 1 and not args.no_recount:
        if rank == 0:
            all_sh = [s for r in range(world) for s in shards_of(r)]
            recount_ok = bool((recount_on_rank0(env, all_sh, parts, mm, fixed, M) == hits).all())
        env.barrier()
    if rank != 0:
        return None
    # parity + CPU baseline: 200 000 of the patterns x a window of this rank's first piece
    parity, cpu = None, None
    if ref.available() and not args.no_cpu_baseline:
        ref.lib()
        g = np.random.default_rng(5)
        sub = np.sort(g.choice(M, min(M, 200_000), replace=False))
        if args.config == "c3":
            skw = dict(concat=pats[sub].reshape(-1), widths=np.full(sub.size, W.C3_WIDTH, np.int32))
        else:
            offs = np.concatenate([[0], np.cumsum(widths, dtype=np.int64)])
            skw = dict(concat=np.concatenate([concat[offs[i]:offs[i + 1]] for i in sub.tolist()]), widths=widths[sub])
        it = pieces.items[0]
        w0 = min(it[1].size, int(2e6))
        sample = it[1][:w0]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pg = bs.PDict(bs.DNAStringSet(**skw), **pkw)
            po = R.PDict(R.SeqSet(**skw), backend="ref", **pkw)
            t0 = time.perf_counter()
            want = R.countPDict(po, sample, max_mismatch=mm, fixed=fixed[0] and fixed[1])
            t_match = time.perf_counter() - t0
            got = bs.countPDict(pg, sample, max_mismatch=mm, fixed=fixed[0] and fixed[1])
        ok = bool((got == want).all())
        parity = {"ok": ok, "sample_letters": int(w0), "sample_patterns": int(sub.size), "sample_matches": int(want.sum())}
        cpu = {"value": w0 / t_match / 1e9 * sub.size / M, "unit": "Gbp/s", "cores": 1, "kind": "reference",
               "sample": f"{sub.size} of the {M} patterns x the first {w0 / 1e6:g} Mbp of rank 0's chunk, {t_match:.2f} s of "
                         "matching; value scaled by the pattern fraction (the Trusted-Band hit rate is proportional to it)",
               "parity_with_gpu_on_sample": ok}
    peak, peak_src = peaks()
    dominant = max(kms, key=kms.get)
    roof = {"bound": "hbm", "achieved": pieces.letters / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": pieces.letters / (ms_per_step * 1e-3) / 1e9 / peak, "traffic": None, "kernel": dominant,
            "peak_source": peak_src, "kernel_ms": kms, "plan": plan,
            "note": "whole step (all kernels of the path) against 1.0 B per subject letter"}
    return {"metric": metric, "value": genome / (ms_per_step * 1e-3) / 1e9, "unit": "Gbp/s", "n_gpus": world,
            "steps": steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": label, "patterns": M, "subject_letters_total": genome,
                       "sharding": f"{world} rank(s), overlap chunks by Trusted-Band end inside every chromosome "
                                   f"(halo {halo[0]} before, {halo[1]} after), {len(pieces.items)} piece(s) on rank 0",
                       "collective": COLLECTIVE_TEXT[COLLECTIVE_USED.get("mode") or args.collective] if world > 1 else "none",
                       "l2": "the rank's chunks are streamed once per step"},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roof,
            "cpu_baseline": cpu, "parity_with_gpu_on_sample": parity,
            "all_reduced_equals_1gpu_recount": recount_ok, "matches_total": int(hits.sum()),
            "index_build": {"pdict_s": t_pdict}}


def run_c4(args, env):
    import biostrings_b200 as bs
    from biostrings_b200 import _lib, matchpdict as mp, sharding, workloads as W
    from oracle import ref, rlevel as R
    torch, L, rank, world = env.torch, env.L, env.rank, env.world
    nreads = max(10_000, int(1e7 * args.scale))
    M = max(1000, int(1e5 * args.scale))
    probes = W.c4_probes(M)
    pdict = bs.PDict(bs.DNAStringSet(concat=probes.reshape(-1), widths=np.full(M, W.C4_WIDTH, np.int32)))
    parts = pdict.parts()
    lo, hi = sharding.shard_bounds(nreads, world, rank)
    host = torch.empty((hi - lo) * W.C4_READ, dtype=torch.uint8, pin_memory=True)
    hv = host.numpy().reshape(hi - lo, W.C4_READ)
    for a in range(lo, hi, 1 << 20):
        b = min(hi, a + (1 << 20))
        hv[a - lo:b - lo] = W.c4_reads(a, b, probes)
    widths = np.full(hi - lo, W.C4_READ, np.int32)
    rset = bs.DNAStringSet(concat=host.numpy(), widths=widths)
    weight = np.ones(hi - lo, dtype=np.int32)
    dsub = mp.DeviceSubject.from_set(rset)
    counts = torch.zeros(M, dtype=torch.int32, device=env.dev)
    letters_rank, letters_total = (hi - lo) * W.C4_READ, nreads * W.C4_READ

    def resident_call():
        c = mp.vmatch_parts(parts, dsub, 0, 0, (True, True), 1, weight, _lib.MATCHES_AS_COUNTS)
        counts.copy_(torch.from_numpy(c))
        env.all_reduce(counts)
        return c

    def host_call():
        with mp.DeviceSubject.from_set(rset) as ds:
            c = mp.vmatch_parts(parts, ds, 0, 0, (True, True), 1, weight, _lib.MATCHES_AS_COUNTS)
        counts.copy_(torch.from_numpy(c))
        env.all_reduce(counts)
        return counts.cpu().numpy()

    resident_call()
    sampler = ClockSampler(env.local)
    sampler.start()
    for _ in range(args.warmup):
        resident_call()
    env.barrier()
    sampler.mark()
    steps = min(args.steps, 20)
    L.bsgpu_kernel_timing(1)
    l0 = L.bsgpu_launch_count()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        resident_call()
    torch.cuda.synchronize()
    te = env.max_over_ranks(time.perf_counter() - t0)
    launches = int(L.bsgpu_launch_count() - l0)
    acc = {}
    for name, ms in _lib.kernel_times():
        acc.setdefault(name, []).append(ms)
    L.bsgpu_kernel_timing(0)
    kms = {k: sum(v) / len(v) for k, v in acc.items()}
    sampler.stop_flag = True
    sampler.join(timeout=2)
    hits = counts.cpu().numpy().astype(np.int64)
    e2e = None
    if not args.no_e2e:
        times = []
        for k in range(3):
            env.barrier()
            t0 = time.perf_counter()
            c_e = host_call()
            if k > 0:
                times.append(time.perf_counter() - t0)
        t_e = env.max_over_ranks(sum(times) / len(times))
        e2e = {"value": letters_total / t_e / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": int(letters_rank + 4 * (hi - lo)),
               "d2h_bytes_per_step": int(4 * M), "ms_per_step": 1e3 * t_e,
               "call": "bsgpu_subject_set(pinned host reads) + bsgpu_vmatch(collapse=1) + all-reduce + counts to the host",
               "counts_equal_device_steps": bool((c_e.astype(np.int64) == hits).all())}
    if rank != 0:
        return None
    parity, cpu = None, None
    if ref.available() and not args.no_cpu_baseline:
        ref.lib()
        k = min(hi - lo, 30_000)
        sub = bs.DNAStringSet(concat=hv[:k].reshape(-1), widths=widths[:k])
        po = R.PDict(R.SeqSet(concat=probes.reshape(-1), widths=np.full(M, W.C4_WIDTH, np.int32)), backend="ref")
        w1 = np.ones(k, dtype=np.int32)
        t0 = time.perf_counter()
        want = np.asarray(R.vcountPDict(po, R.SeqSet(concat=hv[:k].reshape(-1), widths=widths[:k]), collapse=1, weight=w1))
        t_match = time.perf_counter() - t0
        got = np.asarray(bs.vcountPDict(pdict, sub, collapse=1, weight=w1))
        ok = bool(np.array_equal(got, want))
        parity = {"ok": ok, "sample_reads": int(k), "sample_matches": int(want.sum())}
        cpu = {"value": k * W.C4_READ / t_match / 1e9, "unit": "Gbp/s", "cores": 1, "kind": "reference",
               "sample": f"the first {k} reads of rank 0's shard, vcountPDict(collapse=1): {t_match:.2f} s",
               "parity_with_gpu_on_sample": ok}
    peak, peak_src = peaks()
    ms_per_step = 1e3 * te / steps
    dominant = max(kms, key=kms.get) if kms else None
    roof = {"bound": "hbm", "achieved": letters_rank / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": letters_rank / (ms_per_step * 1e-3) / 1e9 / peak, "traffic": None, "kernel": dominant,
            "peak_source": peak_src, "kernel_ms": kms, "plan": L.bsgpu_last_plan().decode(),
            "note": "whole host-API call on the resident set (weights upload, scan, counts to the host) "
                    "against 1.0 B per read letter; wall clock around synchronous calls"}
    return {"metric": "vcountPDict subject Gbp/s (100k 25-mers x 10M reads of 150 bp, collapse=1)",
            "value": letters_total / (ms_per_step * 1e-3) / 1e9, "unit": "Gbp/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"vcountPDict(collapse=1, weight=1L) of {M} 25-mers against {nreads} reads of 150 bp "
                                   "(5 % carry a probe)", "patterns": M, "reads_total": nreads,
                       "sharding": f"{world} rank(s), reads split by sequence, no halo",
                       "collective": "all-reduce of the int32[M] collapsed counts after every call",
                       "l2": "the rank's reads are streamed once per step"},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roof,
            "cpu_baseline": cpu, "parity_with_gpu_on_sample": parity, "matches_total": int(hits.sum())}


def run_ours(args):
    env = Env(args)
    if args.config == "c2":
        line = run_c2(args, env)
    elif args.config == "c4":
        line = run_c4(args, env)
    else:
        line = run_c3_c5(args, env)
    if env.rank == 0:
        print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
