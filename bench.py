#!/usr/bin/env python
"""bench.py -- BN254 G1 MSM at 2^24 points (BASELINE.json's metric) on N B200s of one box, plus NTT 2^24 and the
per-kernel rooflines.

  python bench.py [--gpus N] [--steps K] [--warmup W]                 this repo's CUDA path (N > 1: launched by torchrun)
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] the reference's own CPU (host debug) MSM on the
                                                                      box's host cores, same metric / unit

One JSON line on stdout (rank 0).  A "step" is one MSM of the whole 2^24-point job: the points are sharded by contiguous
range over the N ranks ("scaling": "strong"), every rank runs the full single-GPU pipeline on its slice and the 96-byte
partials are all-gathered (NCCL) and summed.  `value` is device-timed with inputs resident in HBM; `e2e` goes through the
host API (panda_msm_bn254_gpu_with_cached_bases: cached bases, HOST scalars) with the H2D / D2H copies inside the timed
region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 24
MODMUL_MACS = 136            # 2*8^2 + 8 32x32->64 multiply-adds per BN254 Montgomery product
MADD_MODMULS = 10            # XYZZ mixed addition: 8M + 2S
ACC_MODMULS = MADD_MODMULS
ACC_KERNEL = "k_accumulate_range<Bn254> (bucket accumulation, XYZZ mixed additions; one launch per scatter range, pipelined with k_scatter_tiled)"
ACC_KERNEL_BLS = "k_accumulate_range_wide<Bls377> (bucket accumulation, XYZZ mixed additions, 12-limb Fq; one launch per scatter range)"
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # dram__bytes_{read,write}.sum per launch, written by profiles/summarize.py
                                                                        # from the tracked ncu --set full captures


def ncu_traffic(kernel_prefix: str, log_n: int):
    """(bytes over the kernel's launches of ONE MSM, source) of the dominant kernel from the tracked ncu summary, or (None, why)"""
    try:
        data = json.load(open(NCU_TRAFFIC_FILE))
    except Exception as exc:
        return None, f"{os.path.relpath(NCU_TRAFFIC_FILE, ROOT)} unreadable: {exc!r}"
    for entry in data.get("kernels", []):
        if entry["kernel"].startswith(kernel_prefix) and entry.get("log_n") == log_n:
            return (float(entry["dram_bytes_read"]) + float(entry["dram_bytes_write"]),
                    f"{data.get('source', '?')} ({entry['kernel']}, {entry.get('launches', 1)} launches of one MSM summed)")
    return None, f"no capture of {kernel_prefix} at 2^{log_n} in {os.path.relpath(NCU_TRAFFIC_FILE, ROOT)}"


def ntt_passes_count(k: int) -> int:
    return (k + 7) // 8


def kernels_per_msm(plan) -> int:
    """our launches in one device-resident MSM on registered bases: digits, scan x3, (scatter, accumulate) per bucket range, reduce_big,
    bucket_reduce, group_reduce (x2 when there are more than 32 groups), final"""
    return 1 + 3 + 2 * max(1, plan.phases) + 1 + 1 + (2 if plan.groups > 32 else 1) + 1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.thread, self.gpu = [], None, None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path, on the host cores

class RefRunner:
    """concurrent single-threaded calls of the reference's own host (CPU debug) MSM -- oracle/_ref/ref_host_msm, the unmodified
    panda_msm_execute_bn254_host -- one process per host core, all on the same 2^k-point input"""

    def __init__(self, O, tag):
        self.O = O
        self.cores = os.cpu_count() or 1
        self.procs = max(1, min(self.cores, 64))
        self.tmp = os.path.join("/tmp", f"panda_bench_{tag}_{os.getpid()}")
        os.makedirs(self.tmp, exist_ok=True)
        self.files = {}

    def inputs(self, k):
        if k not in self.files:
            n = 1 << k
            fb, fs = os.path.join(self.tmp, f"b{k}.bin"), os.path.join(self.tmp, f"s{k}.bin")
            self.O.gen_bases(0, self.O.seed_for(k), n).tofile(fb)
            self.O.gen_scalars(1, self.O.seed_for(k) + 1, n).tofile(fs)
            self.files[k] = (fb, fs)
        return self.files[k]

    def run(self, k, procs=None):
        """wall seconds of `procs` concurrent reference calls on 2^k points each"""
        procs = procs or self.procs
        fb, fs = self.inputs(k)
        t0 = time.perf_counter()
        ps = [subprocess.Popen([self.O.REF_BIN, fb, fs, str(k), os.path.join(self.tmp, f"o{i}.bin")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
              for i in range(procs)]
        rcs = [p.wait() for p in ps]
        assert all(rc == 0 for rc in rcs), rcs
        return time.perf_counter() - t0

    def fit(self, k_lo=10, k_hi=15):
        """the reference host path costs fixed + per_point * n per call: 2 * 16 * 65535 Jacobian additions of bucket reduction at its
        fixed c = 16 whatever n is (msm_host.cuh:267-370), plus 16 window additions per point.  Two sizes, all cores busy, give both."""
        t_lo, t_hi = self.run(k_lo), self.run(k_hi)
        per_point = max((t_hi - t_lo) / ((1 << k_hi) - (1 << k_lo)), 1e-9)
        fixed = max(t_lo - per_point * (1 << k_lo), 0.0)
        return fixed, per_point, {"fixed_s_per_call": fixed, "per_point_us": per_point * 1e6, "from": f"2^{k_lo}: {t_lo:.2f} s, 2^{k_hi}: {t_hi:.2f} s ({self.procs} concurrent calls)",
                                  "extrapolated_2^24_Mpts_per_s_all_cores": self.procs * (1 << LOG_N) / (fixed + per_point * (1 << LOG_N)) / 1e6,
                                  "note": "extrapolation, not a measurement: every core running one 2^24-point reference call"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import oracle as O     # test infrastructure: allowed here as the reference / cpu_baseline leg only

    kind = "reference" if O.ref_available() else "port"
    budget_s = float(os.environ.get("PANDA_BENCH_REF_BUDGET_S", "300"))
    fit = None
    if kind == "reference":
        rr = RefRunner(O, "ref")
        procs = rr.procs
        fixed, per_point, fit = rr.fit()
        per_step = min(budget_s / max(1, args.steps + args.warmup), 30.0)      # at most 30 s per step however few steps are asked for
        k_sample = 14
        while k_sample < 22 and fixed + per_point * (1 << (k_sample + 1)) <= per_step:
            k_sample += 1
        n_s = 1 << k_sample
        log(f"[bench] reference arm: {procs} processes x 2^{k_sample} points per step (fit: {fixed:.2f} s + {per_point * 1e6:.2f} us/pt per call; "
            f"{fixed / (fixed + per_point * n_s):.0%} of a step is the fixed bucket reduction)")

        def one_step():
            return rr.run(k_sample)
    else:
        procs = max(1, min(os.cpu_count() or 1, 64))
        k_sample = 16
        n_s = 1 << k_sample
        bases = O.gen_bases(0, O.seed_for(LOG_N), n_s)
        scal = O.gen_scalars(1, O.seed_for(LOG_N) + 1, n_s)

        def one_step():
            t0 = time.perf_counter()
            for _ in range(procs):      # the port is internally threaded over windows: run the same number of point-sets
                O.msm(0, bases, scal, n_s, c=16, threads=procs)
            return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    t = [one_step() for _ in range(args.steps)]
    total = sum(t)
    pts = procs * n_s * args.steps
    value = pts / total / 1e6
    fixed_share = (fit["fixed_s_per_call"] / (fit["fixed_s_per_call"] + fit["per_point_us"] * 1e-6 * n_s)) if fit else None
    line = {
        "impl": "reference", "metric": "bn254_g1_msm_2^24_throughput", "value": value, "unit": "Mpts/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 (256-bit Montgomery integers)", "data": "synthetic",
        "config": {"workload": "BN254 G1 MSM, random scalars, reference CPU host path (panda_msm_execute_bn254_host)", "log_n": LOG_N,
                   "sample_points_per_step": procs * n_s, "sample_log_points_per_call": k_sample},
        "cpu_baseline": {"value": value, "unit": "Mpts/s", "cores": procs, "kind": kind,
                         "sample": f"{procs} concurrent single-threaded reference calls of 2^{k_sample} points each per step (the largest power of two that keeps "
                                   f"steps + warm-up within {budget_s:.0f} s); the reference host path is single-threaded with a fixed c = 16, so every call "
                                   f"pays a bucket reduction that does not depend on n" + (f" ({fixed_share:.0%} of this sample's step; it vanishes at 2^24 -- see fit)" if fit else ""),
                         "fit": fit},
        "e2e": {"value": value, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------------------------
# own arm

def cpu_baseline_leg(O, np):
    """reference host path (or the port) on a bounded sample: all cores, 2^16 points per call, plus the fixed / per-point fit."""
    if O.ref_available():
        rr = RefRunner(O, "cpu")
        try:
            fixed, per_point, fit = rr.fit(k_lo=10, k_hi=16)
            k_s = 16
            dt = fixed + per_point * (1 << k_s)          # the 2^16 run of the fit IS the sample
            return {"value": rr.procs * (1 << k_s) / dt / 1e6, "unit": "Mpts/s", "cores": rr.procs, "kind": "reference",
                    "sample": f"{rr.procs} concurrent single-threaded calls of the reference's panda_msm_execute_bn254_host on 2^{k_s} points each ({dt:.1f} s; "
                              f"{fixed / dt:.0%} of it is the n-independent bucket reduction at the reference's fixed c = 16)", "fit": fit}
        except AssertionError:
            pass
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    n_s = 1 << 14
    bases = O.gen_bases(0, 7, n_s); scal = O.gen_scalars(1, 8, n_s)
    t0 = time.perf_counter()
    O.msm(0, bases, scal, n_s, c=16, threads=procs)
    dt = time.perf_counter() - t0
    return {"value": n_s / dt / 1e6, "unit": "Mpts/s", "cores": procs, "kind": "port",
            "sample": f"oracle port (c=16, {procs} threads over windows) on 2^14 points ({dt:.1f} s)"}


def run_multi_capi(ffi, O, np, torch, n, n_dev):
    """2^24-point MSM through panda_msm_execute_bn254_multi_n: ONE process (this rank) drives n_dev GPUs.  Returns the record."""
    from panda_b200.sharded import shard_range

    cfgs = (ffi.MSMConfiguration * n_dev)()
    counts = (C.c_size_t * n_dev)()
    keep, streams, acc = [], [], np.zeros(96, np.uint8)
    for d in range(n_dev):
        lo, hi = shard_range(n, d, n_dev)
        seed = O.seed_for(LOG_N) + 1000 * d
        bh, sh = O.gen_bases(0, seed, hi - lo), O.gen_scalars(1, seed + 1, hi - lo)
        acc = O.jac_add(0, acc, O.expected_progression_msm(0, seed, sh, hi - lo))
        dv = torch.device("cuda", d)
        b, sc, r = torch.from_numpy(bh).to(dv), torch.from_numpy(sh).to(dv), torch.empty(96, dtype=torch.uint8, device=dv)
        assert ffi.lib.panda_set_device(d) == 0
        st = ffi.PandaStream.new()
        assert ffi.lib.panda_msm_register_bases_bn254(b.data_ptr(), hi - lo, st) == 0
        st.sync()
        keep.append((b, sc, r)); streams.append(st)
        counts[d] = hi - lo
        cfgs[d] = ffi.MSMConfiguration(ffi.PandaMemPool.new(d), st, b.data_ptr(), sc.data_ptr(), r.data_ptr(), 0, 0)   # a pool that keeps its memory
    assert ffi.lib.panda_set_device(0) == 0

    def step():
        assert ffi.lib.panda_msm_execute_bn254_multi_n(cfgs, counts, n_dev) == 0

    step()
    streams[0].sync()
    ok = bool((O.jac_to_affine(0, keep[0][2].cpu().numpy()) == O.jac_to_affine(0, acc)).all())
    for _ in range(2):
        step()
    for st in streams:
        st.sync()
    ev0, ev1 = ffi.PandaEvent(None), ffi.PandaEvent(None)
    assert ffi.lib.panda_event_create(C.byref(ev0), True, False) == 0 and ffi.lib.panda_event_create(C.byref(ev1), True, False) == 0
    reps = 5
    t0 = time.perf_counter()
    ev0.record(streams[0])
    for _ in range(reps):
        step()
    ev1.record(streams[0])
    ev1.sync()
    wall_ms = (time.perf_counter() - t0) * 1e3 / reps
    for d in range(n_dev):
        assert ffi.lib.panda_set_device(d) == 0
        streams[d].sync()
        assert ffi.lib.panda_msm_tear_down() == 0
        streams[d].destroy()
    assert ffi.lib.panda_set_device(0) == 0
    del keep
    torch.cuda.empty_cache()
    return {"api": "panda_msm_execute_bn254_multi_n (one process, one host thread per GPU, peer-copied 96-byte partials)", "gpus": n_dev,
            "ms_per_msm_wall": wall_ms, "Mpts_per_s": n / (wall_ms * 1e-3) / 1e6, "verified_against_closed_form": ok,
            "timing": "host wall clock around 5 back-to-back calls bracketed by event syncs on the first GPU's stream (its completion orders all shards)"}


def run_own_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: panda_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")      # host-side barrier for the legs in which ONE rank drives the other ranks' GPUs
    from panda_b200 import gpu_ffi as ffi          # raises if libpanda-cuda.so is missing
    from panda_b200 import gpu_manager as gm
    from panda_b200.sharded import ShardedMsm, shard_range
    import oracle as O                              # checker + cpu_baseline only

    assert ffi.lib.panda_set_device(local_rank) == 0
    dev = torch.device("cuda", local_rank)
    n = 1 << LOG_N
    lo, hi = shard_range(n, rank, world)
    n_local = hi - lo

    # synthetic inputs: rank r owns its own arithmetic progression of points (closed-form expected result per rank)
    seed = O.seed_for(LOG_N) + 1000 * rank
    t0 = time.time()
    bases_h = O.gen_bases(0, seed, n_local)
    scal_h = O.gen_scalars(1, seed + 1, n_local)
    expected_local = O.expected_progression_msm(0, seed, scal_h, n_local)
    log(f"[bench] rank {rank}: inputs for {n_local} points in {time.time() - t0:.1f}s")

    bases_d = torch.from_numpy(bases_h).to(dev)
    scal_d = torch.from_numpy(scal_h).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    pool = ffi.PandaMemPool.new(local_rank)
    sm = ShardedMsm(0)
    # init_msm: the cached bases are announced once; the library builds its table of precomputed multiples here
    assert ffi.lib.panda_msm_register_bases_bn254(bases_d.data_ptr(), n_local, ffi.PandaStream(stream)) == 0
    torch.cuda.synchronize()

    def step():
        return sm.run(bases_d, scal_d, n_local, coord=0, stream=stream, pool=pool.handle)

    # ---- correctness of what is being timed (once, before the timed region)
    res = step()
    torch.cuda.synchronize()
    exp = torch.from_numpy(expected_local).to(dev)
    if world > 1:
        allexp = torch.empty(world * 96, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allexp, exp)
        acc = np.zeros(96, np.uint8)
        for part in allexp.cpu().numpy().reshape(world, 96):
            acc = O.jac_add(0, acc, part)
    else:
        acc = expected_local
    verified = bool((O.jac_to_affine(0, res.cpu().numpy()) == O.jac_to_affine(0, acc)).all())
    if not verified:
        raise SystemExit("bench.py: MSM result does not match the closed form -- refusing to report a number")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed K steps
    # clocks / throttle reasons are sampled from before the warm-up to the end of the timed region (nvidia-smi needs a few hundred
    # milliseconds to deliver its first line, longer than a short timed region); idle samples are dropped by ClockSampler.stop()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while sampler.proc is not None and not sampler.rows and time.time() - t_wait < 2.0:
            time.sleep(0.02)
    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        step()
        marks[i].record()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    per_step = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the host API: cached bases, HOST (pinned) scalars in, 96 bytes out
    mgr = gm.PandaGpuManager.new(local_rank)
    bi = mgr.cache_bases(bases_h)
    pinned = C.c_void_p()
    assert ffi.lib.panda_malloc_host(C.byref(pinned), scal_h.size) == 0
    scal_pinned = np.ctypeslib.as_array((C.c_uint8 * scal_h.size).from_address(pinned.value))
    scal_pinned[:] = scal_h
    out = torch.empty(world * 96, dtype=torch.uint8, device=dev)

    def e2e_step():
        r = gm.panda_msm_bn254_gpu_with_cached_bases(mgr, scal_pinned, bi)     # H2D scalars, MSM, D2H 96 bytes
        if world > 1:
            part = torch.from_numpy(r).to(dev)
            dist.all_gather_into_tensor(out, part)
            assert ffi.lib.panda_msm_combine_bn254(out.data_ptr(), world, out.data_ptr(), 0, ffi.PandaStream(stream)) == 0
            r = out[:96].cpu().numpy()
        return r

    r = e2e_step()
    e2e_ok = bool((O.jac_to_affine(0, r) == O.jac_to_affine(0, acc)).all())
    for _ in range(max(1, args.warmup - 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    ffi.lib.panda_free_host(pinned)
    mgr.deinit()          # panda_msm_tear_down: drops this device's tables, the bench's registered set included
    assert ffi.lib.panda_msm_register_bases_bn254(bases_d.data_ptr(), n_local, ffi.PandaStream(stream)) == 0
    torch.cuda.synchronize()

    del mgr
    extra = {}

    def timed(fn, reps, warm=2):
        """device-timed fn() on torch's current stream: barrier + sync on both sides, max over ranks (ms per call)"""
        for _ in range(warm):
            fn()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        barrier()
        tt = torch.tensor([a0.elapsed_time(a1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def all_ok(flag: bool) -> bool:
        okt = torch.tensor([int(bool(flag))], device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        return bool(okt.item())

    # ---- setup cost of the cached-bases plan and the first-call (windowed) latency: what `value` does not contain
    assert ffi.lib.panda_msm_unregister_bases(bases_d.data_ptr()) == 0
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    assert ffi.lib.panda_msm_register_bases_bn254(bases_d.data_ptr(), n_local, ffi.PandaStream(stream)) == 0
    b1.record()
    torch.cuda.synchronize()
    table_ms, table_bytes = b0.elapsed_time(b1), free0 - torch.cuda.mem_get_info()[0]
    d_tmp = torch.empty(96, dtype=torch.uint8, device=dev)
    wcfg = ffi.MSMConfiguration(pool, ffi.PandaStream(stream), bases_d.data_ptr(), scal_d.data_ptr(), d_tmp.data_ptr(), 0, 0)

    def windowed():          # table_mode 0: the plan an unannounced base pointer gets (no table, W bucket sets, Horner at the end)
        assert ffi.lib.panda_debug_msm_timed(0, wcfg, n_local, 0, 0, 0, None, None) == 0

    first_call_ms = timed(windowed, 2, warm=1)
    extra["setup"] = {"table_build_ms": table_ms, "table_bytes": int(table_bytes), "what": "panda_msm_register_bases_bn254 (init_msm): precomputed multiples "
                      "2^(o_j) * P_i of this rank's points, built once per cached base set; not part of `value` or `e2e`"}
    extra["first_call_ms"] = {"ms": first_call_ms, "what": "same MSM on bases that were never registered: windowed plan, no table, nothing allocated behind the caller "
                              "(per rank, max over ranks)"}

    # ---- config 3's upper end: 2^26 points in total, sharded by point range like the headline job (closed form checked)
    try:
        k26 = 26
        lo26, hi26 = shard_range(1 << k26, rank, world)
        n26 = hi26 - lo26
        seed26 = O.seed_for(k26) + 1000 * rank
        assert ffi.lib.panda_msm_unregister_bases(bases_d.data_ptr()) == 0
        b26 = torch.from_numpy(O.gen_bases(0, seed26, n26)).to(dev)
        s26h = O.gen_scalars(1, seed26 + 1, n26)
        s26 = torch.from_numpy(s26h).to(dev)
        exp26 = torch.from_numpy(O.expected_progression_msm(0, seed26, s26h, n26)).to(dev)
        del s26h
        assert ffi.lib.panda_msm_register_bases_bn254(b26.data_ptr(), n26, ffi.PandaStream(stream)) == 0
        r26 = sm.run(b26, s26, n26, coord=0, stream=stream, pool=pool.handle)
        torch.cuda.synchronize()
        if world > 1:
            allexp26 = torch.empty(world * 96, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allexp26, exp26)
            acc26 = np.zeros(96, np.uint8)
            for part in allexp26.cpu().numpy().reshape(world, 96):
                acc26 = O.jac_add(0, acc26, part)
        else:
            acc26 = exp26.cpu().numpy()
        ok26 = bool((O.jac_to_affine(0, r26.cpu().numpy()) == O.jac_to_affine(0, acc26)).all())
        ms26 = timed(lambda: sm.run(b26, s26, n26, coord=0, stream=stream, pool=pool.handle), 3)
        extra["msm_2^26"] = {"metric": "bn254_g1_msm_2^26_throughput", "ms": ms26, "Mpts_per_s": (1 << k26) / (ms26 * 1e-3) / 1e6, "gpus": world,
                             "points_per_gpu": n26, "verified_against_closed_form": all_ok(ok26)}
        assert ffi.lib.panda_msm_unregister_bases(b26.data_ptr()) == 0
        del b26, s26
        torch.cuda.empty_cache()
    except Exception as exc:      # reported, not hidden
        extra["msm_2^26"] = {"error": repr(exc)[:300]}
    assert ffi.lib.panda_msm_register_bases_bn254(bases_d.data_ptr(), n_local, ffi.PandaStream(stream)) == 0
    torch.cuda.synchronize()

    # ---- N > 1: the sharded four-step NTT at 2^26 (BASELINE.json config 4), NVLink peer stores, device-timed, max over ranks;
    #      checked against the DFT definition at spot indices of every rank's row block (not just a round trip)
    sharded_ntt = None
    if world > 1:
        from panda_b200.sharded_ntt import ShardedNtt, column_block, row_block_indices
        kk = 26
        try:
            os.environ.setdefault("PANDA_ORACLE_THREADS", str(max(1, (os.cpu_count() or 8) // world)))
            sn = ShardedNtt(kk, O.omega_bn254(kk).tobytes(), transport="auto")
            x_nat = O.gen_scalars(1, 26026, 1 << kk)                 # the whole natural-order input on every rank (host): the spot checks need it
            w26 = O.omega_bn254(kk)
            xin = torch.from_numpy(column_block(x_nat, kk, rank, world)).to(dev)
            y = sn.forward(xin)
            torch.cuda.synchronize()
            idx = row_block_indices(kk, rank, world)
            spots = (1, len(idx) // 3 + rank)
            got = y.view(-1, 32)[list(spots)].cpu().numpy()
            ok_dft = all(bool((O.dft_at(1, x_nat, kk, w26, int(idx[p])) == got[i]).all()) for i, p in enumerate(spots))
            del x_nat
            ok_rt = bool((sn.inverse(y) == xin).all().item())
            tms = {"forward": timed(lambda: sn.forward(xin), 5, warm=3)}
            yin = sn.forward(xin).clone()
            tms["inverse"] = timed(lambda: sn.inverse(yin), 5, warm=3)
            xbytes = (1 << kk) // world * 32 * (world - 1) // world
            # where the forward transform's time goes: events between its four steps, mean of 3 calls on this rank, max over ranks per step
            # ("exchange" = the fused twiddle + transpose + peer-store kernel between its two symmetric-memory barriers, so it also holds
            # the wait for the slowest peer)
            phase_ms = {}
            for _ in range(3):
                marks = []
                sn.forward(xin, marks=marks)
                torch.cuda.synchronize()
                for (_, e0), (name, e1) in zip(marks, marks[1:]):
                    phase_ms[name] = phase_ms.get(name, 0.0) + e0.elapsed_time(e1) / 3
            names = sorted(phase_ms)
            pt = torch.tensor([phase_ms[k_] for k_ in names], dtype=torch.float64, device=dev)
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
            phase_ms = {k_: round(float(v), 4) for k_, v in zip(names, pt.tolist())}
            sharded_ntt = {"metric": "bn254_fr_ntt_2^26_sharded_latency", "log_n": kk, "gpus": world, "transport": sn.transport,
                           "forward_ms": tms["forward"], "inverse_ms": tms["inverse"], "dft_spot_checks": all_ok(ok_dft),
                           "spot_indices_per_rank": len(spots), "round_trip_ok": all_ok(ok_rt),
                           "exchange_bytes_per_gpu": xbytes,
                           "phase_ms": phase_ms,
                           "exchange_GBps": ((1 << kk) // world * 32) / (phase_ms["exchange"] * 1e-3) / 1e9 if phase_ms.get("exchange") else None,
                           "exchange_GBps_note": "bytes this GPU's exchange kernel reads (= writes, (N-1)/N of them over NVLink) / the exchange step's time",
                           "layout": "column blocks in, row blocks out (panda_b200/sharded_ntt.py)"}
            del sn, xin, y, yin
            torch.cuda.empty_cache()
        except Exception as exc:      # reported, not hidden
            sharded_ntt = {"error": repr(exc)[:300]}

    # ---- N > 1: the same sharded MSM through the one-process C-ABI entry point (panda_msm_execute_bn254_multi_n: rank 0 drives all N
    #      GPUs from one process, one host thread per GPU; the other ranks have released their memory and wait)
    multi_capi = None
    if world > 1:
        try:
            assert ffi.lib.panda_msm_unregister_bases(bases_d.data_ptr()) == 0
            if rank != 0:
                del bases_d, scal_d
                torch.cuda.empty_cache()
            barrier()
            # the waiting ranks must not sit in an NCCL barrier: its kernel spins on THEIR GPU, which rank 0 is driving here (the two
            # processes' contexts would be time-sliced: measured 39.6 ms per MSM at N = 2 against 18.1 ms for the shards themselves)
            if rank == 0:
                multi_capi = run_multi_capi(ffi, O, np, torch, n, world)
            dist.barrier(group=cpu_group)
        except Exception as exc:
            multi_capi = {"error": repr(exc)[:300]}
            dist.barrier(group=cpu_group)
        if rank == 0:
            assert ffi.lib.panda_set_device(local_rank) == 0
            assert ffi.lib.panda_msm_register_bases_bn254(bases_d.data_ptr(), n_local, ffi.PandaStream(stream)) == 0
            torch.cuda.synchronize()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- per-kernel attribution (rank 0): CUDA events on the launching stream, a separate instrumented pass
    peaks, peak_kind = load_peaks()
    stage = (C.c_float * 7)()
    info = (C.c_uint * 3)()
    d_r = torch.empty(96, dtype=torch.uint8, device=dev)
    cfg = ffi.MSMConfiguration(pool, ffi.PandaStream(stream), bases_d.data_ptr(), scal_d.data_ptr(), d_r.data_ptr(), 0, 0)
    stages = np.zeros(7)
    reps = 3
    for _ in range(reps):
        assert ffi.lib.panda_debug_msm_timed(0, cfg, n_local, 0, 0, -1, stage, info) == 0
        stages += np.array(list(stage))
    stages /= reps
    names = ["digits", "scan", "scatter", "accumulate", "bucket_reduce", "window_reduce", "final"]
    plan = ffi.MsmPlanInfo()
    ffi.lib.panda_debug_msm_plan(0, n_local, info[0], info[1], 0, C.byref(plan))
    W = plan.windows
    entries = n_local * W                                   # mixed additions (zero digits are ~2^-c of them)
    acc_ms = float(stages[3])
    macs = entries * MADD_MODMULS * MODMUL_MACS             # algorithmic 32x32->64 multiply-adds of the accumulate kernel
    peak_ms, peak_ops = C.c_float(), C.c_ulonglong()
    assert ffi.lib.panda_debug_int_peak(1, 4096, C.byref(peak_ms), C.byref(peak_ops)) == 0
    wide_peak = peak_ops.value / (peak_ms.value * 1e-3)
    assert ffi.lib.panda_debug_int_peak(0, 4096, C.byref(peak_ms), C.byref(peak_ops)) == 0
    imad_peak = peak_ops.value / (peak_ms.value * 1e-3)
    assert ffi.lib.panda_debug_int_peak(2, 2048, C.byref(peak_ms), C.byref(peak_ops)) == 0
    modmul_peak = peak_ops.value / (peak_ms.value * 1e-3)
    achieved_macs = macs / (acc_ms * 1e-3)
    acc_bytes = entries * (4 + 64)                          # 4 B sorted index + 64 B gathered affine point per mixed addition
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ranges = max(1, plan.phases)
    digits_bytes = n_local * (32 + 4 * W)                   # scalar read + W 32-bit codes written (grouped by bucket range, tile by tile)
    scatter_bytes = n_local * W * (4 + 4) // ranges         # code read + index written, for the ONE range whose scatter is exposed (stage `scatter`)

    # ---- NTT 2^24 (the second half of BASELINE.json's metric), device-timed
    k = LOG_N
    x = torch.from_numpy(O.gen_scalars(1, 31337, 1 << k)).to(dev)
    y = torch.empty_like(x)
    om = O.omega_bn254(k).copy()
    flag = C.c_uint(0)
    ncfg = ffi.NttconfigurationV1(pool, ffi.PandaStream(stream), x.data_ptr(), y.data_ptr(), om.ctypes.data, k, C.pointer(flag))
    for _ in range(3):
        assert ffi.lib.panda_ntt_execute_bn254_v1(ncfg) == 0
    torch.cuda.synchronize()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record()
    for _ in range(10):
        assert ffi.lib.panda_ntt_execute_bn254_v1(ncfg) == 0
    n1.record()
    torch.cuda.synchronize()
    ntt_ms = n0.elapsed_time(n1) / 10
    # butterflies + pass-boundary twiddles: boundary 1 (2^24 distinct exponents) composes two table entries (2 products),
    # boundary 2 (2^16 distinct) reads a direct table (1 product)
    ntt_modmuls = (1 << k) // 2 * k + (2 + 1) * (1 << k)
    ntt_alg_modmuls = (1 << k) // 2 * k                      # (n/2) log2 n products
    # executed: the last stage of every pass has unit twiddles (no product): 3 passes x 7 stages x n/2 + 3n boundary products
    ntt_exec_modmuls = ntt_passes_count(k) * 0 + ((k - ntt_passes_count(k)) * ((1 << k) // 2)) + 3 * (1 << k)
    ntt_passes = (k + 7) // 8
    pass_ms = (C.c_float * 4)()
    pass_acc = np.zeros(4)
    for _ in range(3):
        assert ffi.lib.panda_debug_ntt_timed(ncfg, 0, pass_ms) == 0
        pass_acc += np.array(list(pass_ms))
    pass_acc /= 3
    # per pass: r_p / 2 butterfly products per element (the last stage of a pass has unit twiddles) + the boundary twiddle
    radices = [k // ntt_passes + (1 if p < k % ntt_passes else 0) for p in range(ntt_passes)]
    ntt_per_pass = []
    for p in range(ntt_passes):
        alg = (1 << k) // 2 * radices[p]
        ntt_per_pass.append({"pass": p, "radix_log": radices[p], "ms": float(pass_acc[p]), "algorithmic_modmul": alg,
                             "frac_of_modmul_peak": alg / (float(pass_acc[p]) * 1e-3) / modmul_peak if pass_acc[p] > 0 else None,
                             "hbm_GBps": 64 * (1 << k) / (float(pass_acc[p]) * 1e-3) / 1e9 if pass_acc[p] > 0 else None})
    del x, y
    torch.cuda.empty_cache()

    # ---- config 5 (one GPU only): BLS12-377 G1 MSM 2^24 with its own roofline block (12-limb Fq: 2 * 12^2 + 12 = 300 MACs per product)
    bls = None
    if world == 1:
        try:
            assert ffi.lib.panda_msm_unregister_bases(bases_d.data_ptr()) == 0
            kb, nbls = LOG_N, 1 << LOG_N
            bb_h = O.gen_bases(1, O.seed_for(kb), nbls)
            sb_h = O.gen_scalars(3, O.seed_for(kb) + 1, nbls)
            exp_b = O.jac_to_affine(1, O.expected_progression_msm(1, O.seed_for(kb), sb_h, nbls))
            bb, sb, rb = torch.from_numpy(bb_h).to(dev), torch.from_numpy(sb_h).to(dev), torch.empty(144, dtype=torch.uint8, device=dev)
            del bb_h, sb_h
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            assert ffi.lib.panda_msm_register_bases_bls12_377(bb.data_ptr(), nbls, ffi.PandaStream(stream)) == 0
            q1.record()
            bcfg = ffi.MSMConfiguration(pool, ffi.PandaStream(stream), bb.data_ptr(), sb.data_ptr(), rb.data_ptr(), kb, 0)

            def bls_step():
                assert ffi.lib.panda_msm_execute_bls12_377(bcfg) == 0

            bls_step()
            torch.cuda.synchronize()
            ok_b = bool((O.jac_to_affine(1, rb.cpu().numpy()) == exp_b).all())
            for _ in range(2):
                bls_step()
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(5):
                bls_step()
            c1.record()
            torch.cuda.synchronize()
            bls_ms = c0.elapsed_time(c1) / 5
            bst, binfo = (C.c_float * 7)(), (C.c_uint * 3)()
            assert ffi.lib.panda_debug_msm_timed(1, bcfg, nbls, 0, 0, -1, bst, binfo) == 0
            assert ffi.lib.panda_debug_msm_timed(1, bcfg, nbls, 0, 0, -1, bst, binfo) == 0
            bplan = ffi.MsmPlanInfo()
            ffi.lib.panda_debug_msm_plan(1, nbls, binfo[0], binfo[1], 0, C.byref(bplan))
            b_entries = nbls * bplan.windows
            b_macs = b_entries * ACC_MODMULS * 300
            b_acc_ms = float(bst[3])
            bls = {"metric": "bls12_377_g1_msm_2^24_throughput", "ms": bls_ms, "Mpts_per_s": nbls / (bls_ms * 1e-3) / 1e6, "verified_against_closed_form": ok_b,
                   "window_bits": bplan.window_bits, "windows": bplan.windows, "table_build_ms": q0.elapsed_time(q1), "table_bytes": int(bplan.table_bytes),
                   "stage_ms": {nm: float(v) for nm, v in zip(names, list(bst))},
                   "roofline": {"kernel": ACC_KERNEL_BLS, "bound": "int32-imad", "achieved": b_macs / (b_acc_ms * 1e-3) / 1e12, "peak": wide_peak / 1e12,
                                "unit": "T(32x32+64 MAC)/s", "frac": b_macs / (b_acc_ms * 1e-3) / wide_peak, "ms": b_acc_ms,
                                "algorithmic": f"{b_entries} bucket additions x {ACC_MODMULS} modmul x 300 MAC (12-limb Fq)",
                                "traffic": ncu_traffic(ACC_KERNEL_BLS.split(" ")[0], LOG_N)[0], "traffic_source": ncu_traffic(ACC_KERNEL_BLS.split(" ")[0], LOG_N)[1]}}
            assert ffi.lib.panda_msm_unregister_bases(bb.data_ptr()) == 0
            del bb, sb, rb
            torch.cuda.empty_cache()
        except Exception as exc:
            bls = {"error": repr(exc)[:300]}

    cpu = cpu_baseline_leg(O, np)
    acc_traffic, acc_traffic_src = ncu_traffic(ACC_KERNEL.split(" ")[0], LOG_N) if n_local == 1 << LOG_N else (None, "per-rank point count differs from the captured 2^24 launch")

    line = {
        "metric": "bn254_g1_msm_2^24_throughput", "value": value, "unit": "Mpts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 (256-bit Montgomery integers)", "data": "synthetic",
        "config": {"workload": "BN254 G1 MSM n=2^24, cached bases + scalars resident in HBM, Jacobian output; points sharded by contiguous range over the ranks",
                   "log_n": LOG_N, "points_per_gpu": n_local, "window_bits": plan.window_bits, "windows": W, "segment_len": plan.segment_len,
                   "parallelism": f"points-shard x{world}", "l2": "inputs (1.5 GiB per 2^24 points) larger than L2; no flush needed",
                   "verified_against_closed_form": verified and e2e_ok},
        "clocks": clocks,
        "e2e": {"value": n / (e2e_ms * 1e-3) / 1e6, "unit": "Mpts/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(scal_h.size) * world,
                "d2h_bytes_per_step": 96 * world, "api": "panda_msm_bn254_gpu_with_cached_bases (host scalars pinned, cached bases) -> panda_msm_execute_bn254_host_scalars: chunked upload overlapped with the sort / accumulation"},
        "gpu_launches": args.steps * world * (kernels_per_msm(plan) + (1 if world > 1 else 0)),
        "roofline": {"kernel": ACC_KERNEL, "bound": "int32-imad", "achieved": achieved_macs / 1e12,
                     "peak": wide_peak / 1e12, "unit": "T(32x32+64 MAC)/s", "frac": achieved_macs / wide_peak,
                     "traffic": acc_traffic, "traffic_source": acc_traffic_src, "algorithmic_bytes": acc_bytes,
                     "peak_source": "IMAD.WIDE issue rate measured live by panda_debug_int_peak (32-bit IMAD rate: %.2f T/s; BN254 modmul microbench: %.2f G modmul/s)"
                                    % (imad_peak / 1e12, modmul_peak / 1e9),
                     "algorithmic": f"{entries} mixed additions x {MADD_MODMULS} modmul x {MODMUL_MACS} MAC (SURVEY 8d counts a squaring as a product; the 2 squarings of "
                                    f"a mixed addition execute 108 MACs each, so frac can read slightly above 1)", "ms": acc_ms,
                     "share_of_step": acc_ms / float(stages.sum())},
        "roofline_hbm": {"bound": "hbm", "kernel": "k_accumulate", "achieved": acc_bytes / (acc_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": acc_bytes / (acc_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
                         "note": "integer-bound kernel: 68 algorithmic bytes per 1360 multiply-adds",
                         "hbm_stages": {"k_digits": {"GB/s": digits_bytes / (float(stages[0]) * 1e-3) / 1e9, "frac": digits_bytes / (float(stages[0]) * 1e-3) / 1e9 / hbm_peak},
                                        "k_scatter": {"GB/s": scatter_bytes / (float(stages[2]) * 1e-3) / 1e9, "frac": scatter_bytes / (float(stages[2]) * 1e-3) / 1e9 / hbm_peak}}},
        "stage_ms": {nm: float(v) for nm, v in zip(names, stages)},
        "stage_note": f"{ranges} bucket ranges: `scatter` is the exposed scatter of the first range, the others run beside the accumulation of the range before "
                      "them and are inside `accumulate` (events on the caller's stream around all range launches)",
        "step_ms_rank0": [round(x, 3) for x in per_step],
        "ntt": {"metric": "bn254_fr_ntt_2^24_latency", "ms": ntt_ms, "passes": ntt_passes, "modmul_per_s": ntt_modmuls / (ntt_ms * 1e-3),
                "frac_of_modmul_peak": ntt_modmuls / (ntt_ms * 1e-3) / modmul_peak, "hbm_GBps": ntt_passes * 64 * (1 << k) / (ntt_ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": ntt_passes * 64 * (1 << k) / (ntt_ms * 1e-3) / 1e9 / hbm_peak,
                "per_pass": ntt_per_pass,
                "roofline": {"bound": "int32-imad", "unit": "G modmul/s", "peak": modmul_peak / 1e9,
                             "achieved": ntt_alg_modmuls / (ntt_ms * 1e-3) / 1e9, "frac": ntt_alg_modmuls / (ntt_ms * 1e-3) / modmul_peak,
                             "algorithmic": "(n/2) log2 n Montgomery products (SURVEY 8d)",
                             "executed_modmul_per_s": ntt_exec_modmuls / (ntt_ms * 1e-3),
                             "executed_frac": ntt_exec_modmuls / (ntt_ms * 1e-3) / modmul_peak,
                             "peak_source": "BN254 Montgomery-product microbenchmark, measured live (panda_debug_int_peak kind 2)"}},
        "cpu_baseline": cpu,
    }
    line.update(extra)
    if bls is not None:
        line["bls12_377"] = bls
    if sharded_ntt is not None:
        line["ntt_sharded"] = sharded_ntt
    if multi_capi is not None:
        line["msm_multi_capi"] = multi_capi
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def emit(line: dict):
    """the ONE JSON line goes to the real stdout; everything else this process (or NCCL) prints was redirected to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)


def main():
    os.dup2(2, 1)          # libraries that print to stdout (NCCL's version banner) must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.warmup < 3:
        log("[bench] note: timing rules ask for >= 3 warm-up steps")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", str(29400 + os.getpid() % 500), os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
                   "--warmup", str(args.warmup)]
            return subprocess.call(cmd, stdout=_REAL_STDOUT)
        log(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    return run_own_arm(args)


if __name__ == "__main__":
    sys.exit(main())
