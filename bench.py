#!/usr/bin/env python
"""bench.py -- countPDict subject Gbp/s (BASELINE.json metric) on N B200s of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mm 0|2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): countPDict of 1M 25-mers (C2 dictionary: random, 10 % cut from
the subject, 1 % duplicates) against a synthetic chr1-shaped subject of 250 Mbp per GPU
(i.i.d. letters, GC 0.41, ~7 % N in 40 runs).  With N GPUs the subject is N x 250 Mbp, split
by Trusted-Band end with an overlap of width-1 (weak scaling); the per-pattern counts are
all-reduced over NCCL inside the timed region.

A step = one pass of the hot path over the rank's resident chunk of 1-byte letters, treated
as a new batch (derived planes dropped): scan (max.mismatch 0: the byte-seed scan reads the
letters directly; max.mismatch 2: pack to 2-bit planes, q-gram scan), counting
into the rank's count vector, which accumulates over the batches.  `value` times EXACTLY K
steps between barrier + synchronize, one CUDA event on each side on the launching stream, max
over ranks; the path shards without a data-path collective, its one exchange (the all-reduce
of the accumulated counts) runs once after the last step, inside the bracket.  No L2 flush
inside the bracket: every step streams a
250 MB subject through the 126 MB L2, so nothing of it survives to the next step; the
dictionary index staying in L2 is the steady state of the path.  `e2e` times the public
host-buffer call (pinned host bytes -> H2D -> scan -> D2H counts).  `cpu_baseline` runs the
reference's own C (oracle/_ref, else our C restatement) on one host core over a bounded
sample of the same workload and doubles as a parity check of that sample.
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
                    help="max.mismatch of the headline metric (0 = exact, 2 = MTB_PDict)")
    ap.add_argument("--subject-mbp", type=float, default=250.0, help="letters per GPU, in Mbp")
    ap.add_argument("--patterns", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-mbp", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mm2", action="store_true")
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
    kind, _, _ = cpu_backend()
    cores = max(1, min(os.cpu_count() or 1, 32))
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
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": "Gbp/s", "cores": cores, "kind": kind,
                             "sample": f"{cores} x {per_core / 1e6:g} Mbp slices of the subject per step, "
                                       "one single-threaded process per core, ACtree2 index built once per "
                                       "process outside the timed part"},
            "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"countPDict, {args.patterns} x {WIDTH}-mers (Trusted Band = whole pattern, "
                        f"max.mismatch={args.mm}), synthetic chr1-shaped subject of "
                        f"{args.subject_mbp:g} Mbp per GPU ({args.subject_mbp * world:g} Mbp total)",
            "patterns": args.patterns, "pattern_width": WIDTH, "max_mismatch": args.mm,
            "subject_letters_per_gpu": int(args.subject_mbp * 1e6),
            "sharding": f"{world} overlap chunk(s) by Trusted-Band end, halo = width-1",
            "l2": "inputs larger than L2: every step streams the 250 MB subject chunk through the 126 MB L2 "
                  "(no flush inside the bracket)",
            "step": "new batch of 1-byte letters resident in HBM -> scan into the rank's accumulating count vector "
                    "(max.mismatch 0: the byte-seed scan reads the letters directly; otherwise pack to 2-bit "
                    "planes first); one all-reduce of the counts after the last step, inside the timed region"}


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(local):
    """Best effort: run this rank on the CPUs of its GPU's NUMA node BEFORE it allocates its pinned
    host buffer, so that the buffer the end-to-end leg streams over PCIe lives in the memory next
    to that GPU's root port instead of crossing the inter-socket link (with 8 ranks per box the
    e2e leg is one host-memory -> PCIe stream per GPU).  Returns what was done (for `config`), or
    None when the topology cannot be read.  BENCH_NUMA=0 disables it."""
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    import biostrings_b200 as bs
    from biostrings_b200 import _lib, matchpdict as mp, sharding, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the PDict path has no CPU fallback)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    L = _lib.lib()
    _lib.check(L.bsgpu_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)

    chrom_len = int(args.subject_mbp * 1e6)
    subject_len = chrom_len * world
    M = args.patterns
    pats = synth.dictionary(M, WIDTH, chrom_len, world)
    t_build0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pdict = bs.PDict(bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(M, WIDTH, np.int32)),
                         max_mismatch=None if args.mm == 0 else args.mm)
    t_pdict = time.perf_counter() - t_build0
    parts = pdict.parts()
    parr = mp._parts_array(parts)
    halo = sharding.halo_of(pdict)
    lo, hi = sharding.shard_bounds(subject_len, world, rank)
    a, b = sharding.chunk_range(lo, hi, subject_len, halo[0], halo[1])
    host = torch.empty(b - a, dtype=torch.uint8, pin_memory=True)
    chunk = host.numpy()
    synth.subject_region(a, b, chrom_len, out=chunk)
    dsubj = mp.DeviceSubject.from_chunk(chunk, a, subject_len, lo, hi)
    counts = torch.zeros(M, dtype=torch.int32, device=dev)
    nl = C.c_int(0)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_steps(parr_, nparts_, mm_, nsteps):
        """EXACTLY nsteps steps bracketed by barrier + synchronize and one CUDA event on each side
        of the launching stream; max over ranks.  A step = a new batch of letters (derived planes
        dropped) counted into the rank's count vector, which accumulates over the batches as it
        does over the chunks of a long subject; the path shards without a data-path collective
        (SURVEY 8e), so the one exchange -- the all-reduce of the accumulated counts -- happens
        once, after the last step, inside the bracket.  The library times its own kernels with
        events on the launching stream while this runs.
        Returns (total_ms, {kernel: mean ms}, launches, matches per step)."""
        L.bsgpu_kernel_timing(1)
        l0 = L.bsgpu_launch_count()
        barrier()
        ev0.record(stream)
        counts.zero_()
        for k in range(nsteps):
            _lib.check(L.bsgpu_subject_repack(dsubj.handle, sptr))
            _lib.check(L.bsgpu_count_device(parr_, nparts_, dsubj.handle, mm_, 0, 1, 1,
                                            C.c_void_p(counts.data_ptr()), sptr, C.byref(nl)))
        if world > 1:
            dist.all_reduce(counts)
        ev1.record(stream)
        barrier()
        launches_ = int(L.bsgpu_launch_count() - l0)
        acc = {}
        for name, ms in _lib.kernel_times():
            acc.setdefault(name, []).append(ms)
        L.bsgpu_kernel_timing(0)
        tt = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total = int(counts.sum(dtype=torch.int64).item())
        assert total % nsteps == 0, "every step must find the same matches"
        return tt.item(), {k: sum(v) / len(v) for k, v in acc.items()}, launches_, total // nsteps

    def step():
        counts.zero_()
        _lib.check(L.bsgpu_subject_repack(dsubj.handle, sptr))
        _lib.check(L.bsgpu_count_device(parr, len(parts), dsubj.handle, args.mm, 0, 1, 1,
                                        C.c_void_p(counts.data_ptr()), sptr, C.byref(nl)))
        if world > 1:
            dist.all_reduce(counts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # first call: builds the device index of the dictionary (cached in the PDict3Parts objects)
    torch.cuda.synchronize()
    t_first0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t_first0
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.mark()
    total_ms, kms, launches, total_hits = timed_steps(parr, len(parts), args.mm, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    plan = L.bsgpu_last_plan().decode()
    ms_per_step = total_ms / args.steps
    letters_per_rank = hi - lo
    letters_per_rank_pre = letters_per_rank
    value = world * letters_per_rank / (ms_per_step * 1e-3) / 1e9

    # ---- the other half of the metric: max.mismatch = 2 on the same dictionary (MTB_PDict) ----
    mm2 = None
    if args.mm == 0 and not args.no_mm2:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pd2 = bs.PDict(bs.DNAStringSet(concat=pats.reshape(-1), widths=np.full(M, WIDTH, np.int32)),
                           max_mismatch=2)
        parts2 = pd2.parts()
        parr2 = mp._parts_array(parts2)
        n2 = max(3, min(args.steps, 10))
        for _ in range(2):
            counts.zero_()
            _lib.check(L.bsgpu_count_device(parr2, len(parts2), dsubj.handle, 2, 0, 1, 1,
                                            C.c_void_p(counts.data_ptr()), sptr, C.byref(nl)))
        barrier()
        tot2, kms2, l2, hits2 = timed_steps(parr2, len(parts2), 2, n2)
        mm2 = {"metric": "countPDict subject Gbp/s (1M 25-mers, max.mismatch 2)",
               "value": world * letters_per_rank_pre / (tot2 / n2 * 1e-3) / 1e9, "unit": "Gbp/s",
               "steps": n2, "ms_per_step": tot2 / n2, "kernel_ms": kms2,
               "gpu_launches": l2, "matches_total": hits2,
               "plan": L.bsgpu_last_plan().decode()}
        del pd2, parts2, parr2
        counts.zero_()
        _lib.check(L.bsgpu_count_device(parr, len(parts), dsubj.handle, args.mm, 0, 1, 1,
                                        C.c_void_p(counts.data_ptr()), sptr, C.byref(nl)))
        if world > 1:
            dist.all_reduce(counts)
        barrier()

    # ---- matchPDict (MATCHES_AS_ENDS) on the resident chunk: hits as a CSR on the host ------
    ends_leg = None
    if args.mm == 0 and not args.no_e2e:
        times = []
        for k in range(4):
            barrier()
            t0 = time.perf_counter()
            off_e, ends_e = mp.match_parts(parts, dsubj, 0, 0, (True, True), _lib.MATCHES_AS_ENDS)
            torch.cuda.synchronize()
            if k > 0:
                times.append(time.perf_counter() - t0)
        te = sum(times) / len(times)
        ends_leg = {"call": "matchPDict on the resident chunk (records -> radix sort -> CSR on the host)",
                    "ms": 1e3 * te, "hits_per_rank": int(ends_e.size), "Mhits_per_s": ends_e.size / te / 1e6,
                    "Gbp_s_per_gpu": letters_per_rank / te / 1e9}

    # ---- end to end through the public host-buffer call -----------------------------------
    e2e = None
    if not args.no_e2e:
        n_e2e = max(2, min(args.steps, 5))
        times = []
        for k in range(n_e2e + 1):
            barrier()
            t0 = time.perf_counter()
            with mp.DeviceSubject.from_chunk(chunk, a, subject_len, lo, hi) as ds:      # H2D + pack
                c = mp.match_parts(parts, ds, args.mm, 0, (True, True), _lib.MATCHES_AS_COUNTS)  # + D2H
            if world > 1:
                ct = torch.from_numpy(c).to(dev)
                dist.all_reduce(ct)
                c = ct.cpu().numpy()
            torch.cuda.synchronize()
            if k > 0:
                times.append(time.perf_counter() - t0)
            if os.environ.get("BENCH_DEBUG"):
                print(f"[e2e {k}] {(time.perf_counter() - t0) * 1e3:.2f} ms", file=sys.stderr)
        te = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * letters_per_rank / te.item() / 1e9, "unit": "Gbp/s",
               "h2d_bytes_per_step": int(b - a), "d2h_bytes_per_step": int(4 * M),
               "ms_per_step": 1e3 * te.item(),
               "call": "DeviceSubject.from_chunk(pinned host bytes) + countPDict (bsgpu_subject_chunk + "
                       "bsgpu_match through the C ABI), counts read back to the host",
               "host_binding": numa if numa is not None else "none (topology not readable or BENCH_NUMA=0)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peak, peak_src = peaks()
    dominant = max(kms, key=kms.get)
    dom_ms = kms[dominant]
    scan_ms = dom_ms
    alg_bytes = float(letters_per_rank)            # 1.0 B per subject letter
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dominant, {}).get("dram_bytes_per_launch")
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "kernel": dominant,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
            "kernel_ms": kms,
            "kernel_ms_source": "CUDA events recorded by the library around each of its launches, on the "
                                "launching stream, inside the timed region (bsgpu_kernel_timing); mean per launch",
            "step_frac_of_hbm_peak": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
            "plan": plan,
            "note": "algorithmic bytes = 1.0 B per subject letter (the reference's 1-byte encoding, SURVEY 8d); "
                    "the exact scan reads the letters once (TMA) and is bound by L2 random-sector gathers "
                    "(one 16 B Bloom block per probe, one probe every S letters), see l2_gather"}
    probe = os.path.join(ROOT, "profiles", "r01_probe_peaks.json")
    if os.path.exists(probe):
        with open(probe) as f:
            pk = json.load(f)
        g = pk.get("gather_Gsectors_64MB") or pk.get("gather_Gsectors_32MB")
        if g:
            stride_s = 1
            for tok in plan.split(" "):
                if tok.startswith("S="):
                    stride_s = int(tok[2:])
            rate = letters_per_rank / stride_s / (scan_ms * 1e-3) / 1e9
            roof["l2_gather"] = {"achieved_Gprobes_s": rate, "peak_Gsectors_s": g, "frac": rate / g,
                                 "probes_per_letter": 1.0 / stride_s,
                                 "peak_source": "tools/probe_peaks.py on this pool (profiles/r01_probe_peaks.json)"}

    # ---- CPU baseline on a bounded sample (also a parity check of that sample) --------------------
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline and args.mm == 0:
        kind, PPTB, mod = cpu_backend()
        n = int(min(args.cpu_sample_mbp * 1e6, chrom_len - 20_000))
        s_lo = 20_000
        sample = synth.subject_region(s_lo, s_lo + n, chrom_len)
        want, t_build, t_match = cpu_count_exact(PPTB, mod, pats, sample)
        got = bs.countPDict(pdict, sample)
        parity = bool((got == want).all())
        cpu = {"value": n / t_match / 1e9, "unit": "Gbp/s", "cores": 1, "kind": kind,
               "sample": f"first {n / 1e6:g} Mbp after the telomere of the same subject, same 1M-pattern "
                         f"dictionary; ACtree2 index build {t_build:.2f} s not included, cold tree; "
                         f"{t_match:.2f} s of matching",
               "parity_with_gpu_on_sample": parity, "sample_matches": int(want.sum())}

    line = {"metric": f"countPDict subject Gbp/s (1M 25-mers, max.mismatch {args.mm})",
            "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, world), "clocks": sampler.summary(),
            "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "mm2": mm2, "matches_total": total_hits, "match_ends": ends_leg,
            "index_build": {"pdict_s": t_pdict, "first_call_s": t_first,
                            "note": "PDict() = host preprocessing (Trusted-Band hash, duplicates) + upload; the "
                                    "first call also builds and uploads the engine's own index (byte-seed table "
                                    "and Bloom filter), cached in the PDict3Parts objects"}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
