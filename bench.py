#!/usr/bin/env python
"""Benchmark of the exhaustive LCC search (BASELINE.json metric: rotations/s at 128^3,
10 degree search, Laplace pre-filter -- configs[1]) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" scans one block of `--rot-per-step` rotations per GPU (default at 128^3: the whole 7416-rotation
10 degree search; weak scaling: every rank
gets its own block of that size) into the device-resident packed best grid; at N > 1 the
step ends with the packed MAX all-reduce that merges the ranks' grids.  `value` is
rotations/s over all ranks with inputs resident in HBM; `e2e` is the same metric through
the C-ABI call that takes HOST buffers (`pfb_search_host`: uploads, FT(map), search,
unpack, downloads).  `roofline` follows SURVEY.md section 8(d): B_rot = (2 n_f + 6) S bytes per
rotation.  `cpu_baseline` times the reference CPU path (oracle port; rotation by the reference's own compiled
_extensions.c when oracle/_ref travelled) on this box's cores on a bounded rotation sample: one persistent
process per core, correlator built once per process outside the timed region, >= 8 rotations per process.
At N > 1 the line also carries `merge_check` (the NCCL-merged grids of a sharded search equal the single-GPU
grids bit for bit and the reference golden) and, at every N, `strong`: ONE 7416-rotation core-weighted search
(configs[2]) sharded over the N ranks through CUDACorrelator.scan().  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (synth config, laplace, core_weighted, n, search angle label, rotations of the full search)
    "config2": dict(n=128, laplace=True, cw=False, angle="10deg", full_R=7416,
                    desc="GroEL-sized synthetic 128^3 map @23.5A, GroES-sized template, 10deg search (7416 rot), Laplace"),
    "config3": dict(n=128, laplace=False, cw=True, angle="10deg", full_R=7416,
                    desc="synthetic 128^3 map, 10deg search, core-weighted LCC"),
    "config1": dict(n=64, laplace=False, cw=False, angle="20deg", full_R=648,
                    desc="synthetic 64^3 map @8A, 300-residue model, 20deg search (648 rot)"),
    "config5": dict(n=192, laplace=False, cw=False, angle="2.5deg", full_R=207576,
                    desc="synthetic 192^3 map @8A, one of 4 sub-unit templates, 2.5deg search, plain LCC"),
    "config4": dict(n=256, laplace=True, cw=True, angle="4.71deg", full_R=70728,
                    desc="ribosome-sized synthetic 256^3 map @6A, 4.71deg search, Laplace + core-weighted"),
    # not a BASELINE config: a non-cubic grid of the kind the reference CLI's trim + extend produces, on the per-axis
    # fused pipeline (the input of tests/golden/scan_96x128x64_laplace.npz)
    "cli_96x128x64": dict(n=128, shape=(96, 128, 64), laplace=True, cw=False, angle="20deg", full_R=648,
                          desc="synthetic 96x128x64 map @9A (a CLI-like trimmed + extended grid), Laplace"),
}


def shape_of(w):
    return tuple(w.get("shape", (w["n"],) * 3))


def spectrum_bytes(w):
    """S of SURVEY 8(d): one complex64 half-spectrum."""
    nz, ny, nx = shape_of(w)
    return 8 * nz * ny * (nx // 2 + 1)


def make_inputs(workload):
    from powerfit_b200 import synth
    w = WORKLOADS[workload]
    if workload == "config1":
        case = synth.config1(seed=0)
    elif workload == "config4":
        case = synth.config4(seed=0)
    elif workload == "config5":
        case = synth.config5(seed=0)
    elif workload == "cli_96x128x64":
        case = synth.make_case(shape=(96, 128, 64), voxelspacing=3.0, resolution=9.0, n_res=150, rg=11.0, n_copies=3, seed=31)
    else:
        case = synth.config2(seed=0, core_weighted=w["cw"])
    return case


def search_rotations(workload, count):
    """`count` rotation matrices of the workload's search.  The 128^3 workloads use the reference's own 10 degree
    set (7416 proportional orientations, stored with the full-search golden made by the real reference), repeated
    when a run needs more; the other sizes have no stored set here and use seeded uniform random rotations."""
    from powerfit_b200 import synth
    path = os.path.join(ROOT, "tests", "golden", "scan_config2_128_full.npz")
    if WORKLOADS[workload]["angle"] == "10deg" and os.path.exists(path):
        full = np.load(path)["rotations"]
        reps = (count + len(full) - 1) // len(full)
        return np.ascontiguousarray(np.tile(full, (reps, 1, 1))[:count]), "c48 10deg set (7416 proportional orientations)"
    return synth.random_rotations(count, seed=1), "seeded uniform random rotations"


def default_batch(w):
    """pfb_plan_create's default rotations per batch (csrc/api.cu)."""
    per_pair = 6 * int(np.prod(shape_of(w))) * 8
    pairs = max(1, min(256, (32768 << 20) // per_pair))
    return 2 * pairs


def workload_config(w, world, rps, batch, rot_desc):
    """The `config` object both arms print (the reference arm times a bounded sample of this workload)."""
    nf = 3 if w["cw"] else 2
    return {"workload": w["desc"], "shape": list(shape_of(w)), "rotations_per_step_per_gpu": int(rps), "rotation_set": rot_desc,
            "batch": int(batch), "mask": "core-weighted" if w["cw"] else "binary", "laplace": w["laplace"],
            "parallelism": "rotation shards x%d + packed MAX all-reduce" % world,
            "l2": "working set per batch (%.0f MB) exceeds the 126 MB L2; no flush needed"
                  % (batch / 2 * (nf + 3) * 8 * int(np.prod(shape_of(w))) / 1e6)}


def default_rps(w):
    n = w["n"]
    if "shape" in w:
        return 4096
    # one step = one full rotational search of the named angle at 128^3 (7416 rotations); bounded blocks elsewhere
    return {64: 2048, 128: w["full_R"], 256: 512, 192: 808}.get(n, 256)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm
_CPU = {}


def _cpu_init(target, template, mask, laplace, root):
    """Pool initialiser: one correlator per process, built once (the reference's _cpu_scan builds one per
    forked process too, powerfitter.py:110-122); rotation through the reference's own compiled C when it is there."""
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import oracle as O
    ext = O.load_reference_extension()
    rotate = None
    if ext is not None:
        def rotate(grid, rotmat, radius, out, nearest):
            ext.rotate_grid3d(grid, np.ascontiguousarray(rotmat, dtype=np.float64), int(radius), out, bool(nearest))
    c = O.OracleCorrelator(target, laplace=laplace, rotate=rotate)
    c.template = template
    c.mask = mask
    _CPU["c"] = c


def _cpu_block(rotations):
    c = _CPU["c"]
    c.rotations = rotations
    c.scan()
    return c.lcc, c.rot


class CpuArm:
    """The reference CPU search on this box's cores: `cores` persistent processes, contiguous rotation blocks
    (powerfitter.py:95-108), strict-'>' merge in block order (:146-163)."""

    def __init__(self, case, laplace, cores=None):
        import multiprocessing as mp
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=False)
        from oracle import oracle as O
        self.O = O
        self.kind = "reference" if O.load_reference_extension() is not None else "port"
        self.cores = cores or os.cpu_count() or 1
        self.shape = case.target.shape
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init,
                                                initargs=(case.target, case.template, case.mask, laplace, ROOT))
        self.pool.map(_cpu_block, [np.eye(3)[None]] * self.cores)       # every process has built its correlator

    def step(self, rotations):
        blocks = self.O.partition_rotations(len(rotations), self.cores)
        t0 = time.perf_counter()
        parts = self.pool.map(_cpu_block, [rotations[a:b] for a, b in blocks], chunksize=1)
        self.O.combine_partials(parts, len(rotations) // self.cores, self.shape)
        return time.perf_counter() - t0

    def describe(self, n, per_core):
        how = ("rotate_grid3d = the reference's own _extensions.c (oracle/_ref), FFTs = numpy.fft (the reference's "
               "fallback without pyFFTW), conj_multiply/calc_lcc restated") if self.kind == "reference" else \
              "oracle port of CPUCorrelator (numpy.fft backend)"
        return ("%d rotations per step (%d per process x %d persistent processes, correlator and FT(map) built once "
                "per process outside the timed region), FP64, %s" % (n, per_core, self.cores, how))

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_per_core(n):
    """Rotations per process and step: >= 8, more on small grids; ~10-30 s of CPU work in total."""
    return 8 if n >= 128 else 48


def cpu_baseline(case, laplace, rotations, n):
    """Reference CPU path on a bounded sample of the same workload (one warm-up step, one timed step)."""
    arm = CpuArm(case, laplace)
    try:
        per_core = cpu_per_core(n)
        cnt = arm.cores * per_core
        arm.step(rotations[:arm.cores])                    # warm-up: FFT plans, page faults
        dt = arm.step(rotations[:cnt])
        return {"value": cnt / dt, "unit": "rotations/s", "cores": arm.cores, "kind": arm.kind,
                "sample": arm.describe(cnt, per_core) + ", %.1f s" % dt}
    finally:
        arm.close()


def run_reference(args, w, rank, world, emit):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores (the reference's
    own Python package cannot travel; see CpuArm), on a bounded sample of our arm's workload per step."""
    if rank != 0:
        return
    case = make_inputs(args.workload)
    n = w["n"]
    arm = CpuArm(case, w["laplace"])
    per_core = cpu_per_core(n)
    cnt = arm.cores * per_core
    rots, rot_desc = search_rotations(args.workload, cnt * (args.steps + args.warmup))
    times = []
    for s in range(args.warmup + args.steps):
        dt = arm.step(rots[s * cnt:(s + 1) * cnt])
        if s >= args.warmup:
            times.append(dt)
    arm.close()
    T = sum(times)
    val = cnt * args.steps / T
    sample = arm.describe(cnt, per_core)
    out = {"impl": "reference", "metric": "rotations/s (LCC search)", "value": val, "unit": "rotations/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(w, world, args.rot_per_step or default_rps(w), args.batch or default_batch(w), rot_desc),
           "cpu_baseline": {"value": val, "unit": "rotations/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
           "e2e": {"value": val, "unit": "rotations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


# ----------------------------------------------------------------------------- multi-GPU checks
def merge_check(dev, rank, world):
    """Sharded CUDACorrelator.scan() (rotation blocks per rank + ONE NCCL MAX all-reduce on the packed keys) against
    the unsharded scan on the same rank and against the reference golden.  (a) tests/golden/scan_32_plain.npz, the
    any-shape pipeline, 216 rotations: golden tolerance (1e-4, rotation index where decided) and agreement with the
    unsharded result; (b) 64 rotations of the 128^3 core-weighted workload on the fused pipeline: bit for bit."""
    import torch
    import torch.distributed as dist
    from powerfit_b200 import CUDACorrelator, synth
    res = {}
    g = np.load(os.path.join(ROOT, "tests", "golden", "scan_32_plain.npz"))
    tgt, tmpl, msk = (g[k].astype(np.float64) for k in ("target", "template", "mask"))
    out = {}
    for shard in (True, False):
        c = CUDACorrelator(tgt, device=dev, laplace=bool(g["laplace"]), shard=shard)
        c.template, c.mask, c.rotations = tmpl, msk, g["rotations"]
        c.scan()
        out[shard] = (c.lcc.copy(), c.rot.copy())
        del c
    decided = (g["lcc"] - g["lcc2"]) > 1e-4
    err = float(np.abs(out[True][0] - g["lcc"]).max())
    res["golden_32"] = {"max_abs_dlcc": err, "rot_equal_where_decided": bool(np.array_equal(out[True][1][decided], g["rot"][decided])),
                        "sharded_vs_single_max_abs": float(np.abs(out[True][0] - out[False][0]).max()),
                        "sharded_vs_single_rot_equal_where_decided": bool(np.array_equal(out[True][1][decided], out[False][1][decided]))}
    ok = err <= 1e-4 and res["golden_32"]["rot_equal_where_decided"] and \
        res["golden_32"]["sharded_vs_single_max_abs"] <= 1e-5 and res["golden_32"]["sharded_vs_single_rot_equal_where_decided"]
    case = synth.config2(seed=0, core_weighted=True)
    rots, _ = search_rotations("config3", 7416)
    sub = np.ascontiguousarray(rots[::7416 // 64][:64])
    out = {}
    for shard in (True, False):
        c = CUDACorrelator(case.target, device=dev, laplace=False, shard=shard)
        c.template, c.mask, c.rotations = case.template, case.mask, sub
        c.scan()
        out[shard] = (c.lcc.copy(), c.rot.copy())
        fused = c.plan_info(6)
        del c
    same = bool(np.array_equal(out[True][0], out[False][0]) and np.array_equal(out[True][1], out[False][1]))
    res["fused_128_64rot"] = {"bit_identical": same, "fused": int(fused), "nonzero": int((out[True][0] > 0).sum())}
    ok = ok and same and res["fused_128_64rot"]["nonzero"] > 0
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["ok"] = bool(flag.item())
    res["ranks"] = world
    return res


def strong_search(dev, rank, world, batch):
    """ONE 10 degree search (7416 rotations) of the 128^3 core-weighted workload (BASELINE configs[2]) sharded over
    the ranks through CUDACorrelator.scan(): contiguous blocks (powerfitter.py:95-108), packed MAX all-reduce,
    unpack, download on every rank.  Correlator set-up (upload, FT(map), template preparation) is outside the
    timed call, like the reference CLI's "Time for search".  Times are the max over ranks."""
    import torch
    import torch.distributed as dist
    from powerfit_b200 import CUDACorrelator, synth
    case = synth.config2(seed=0, core_weighted=True)
    rots, rot_desc = search_rotations("config3", 7416)
    c = CUDACorrelator(case.target, device=dev, laplace=False, batch=batch, shard=True)
    c.template, c.mask, c.rotations = case.template, case.mask, rots
    c.scan()                                             # warm-up (NCCL channels, first-use set-up)
    best = None
    for rep in range(3):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        c.scan()
        dt = time.perf_counter() - t0
        prof = dict(c.last_scan_profile)
        vals = torch.tensor([dt, prof["search_ms"], prof["allreduce_ms"], prof["unpack_download_ms"]], device=dev,
                            dtype=torch.float64)
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        vals = [float(v) for v in vals.tolist()]
        if best is None or vals[0] < best[0]:
            best = vals
    checksum = int(np.count_nonzero(c.lcc)), float(c.lcc.max()), int(c.rot.reshape(-1)[int(np.argmax(c.lcc))])
    # the same search with the merged grids delivered to rank 0 only (MAX reduce, one download): what the reference's
    # multi-process search does -- the workers hand their partial grids to the parent (powerfitter.py:135-163)
    root_best = None
    if world > 1:
        c.result_rank = 0
        c.scan()
        for rep in range(3):
            torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            c.scan()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            root_best = float(dt.item()) if root_best is None else min(root_best, float(dt.item()))
        if rank == 0:
            assert int(np.count_nonzero(c.lcc)) == checksum[0] and float(c.lcc.max()) == checksum[1]
        c.result_rank = None
    extra = {} if root_best is None else {"seconds_result_on_rank0": root_best,
                                          "rotations_per_s_result_on_rank0": 7416 / root_best}
    return {**extra, "workload": "ONE 7416-rotation search, 128^3 core-weighted (configs[2]), sharded over %d rank(s)" % world,
            "rotation_set": rot_desc, "rotations": 7416, "rotations_per_rank": int(c.last_scan_profile["rotations"]),
            "seconds": best[0], "rotations_per_s": 7416 / best[0], "search_ms": best[1], "allreduce_ms": best[2],
            "unpack_download_ms": best[3], "timing": "wall clock of scan() between barriers, max over ranks, best of 3; "
            "split from CUDA events on the scan stream",
            "result": {"nonzero": checksum[0], "lcc_max": checksum[1], "rot_at_max": checksum[2]}}


def sharded_config4(dev, rank, world, R=1024):
    """BASELINE configs[3]: the 256^3 Laplace + core-weighted workload, ONE block of R rotations of the search sharded
    over the ranks through CUDACorrelator.scan() (contiguous blocks, packed MAX all-reduce, unpack + download on every
    rank).  Set-up outside the timed call; wall clock between barriers, max over ranks, best of 2."""
    import torch
    import torch.distributed as dist
    from powerfit_b200 import CUDACorrelator
    w = WORKLOADS["config4"]
    case = make_inputs("config4")
    rots, rot_desc = search_rotations("config4", R)
    c = CUDACorrelator(case.target, device=dev, laplace=w["laplace"], shard=True)
    c.template, c.mask, c.rotations = case.template, case.mask, rots
    c.scan()                                             # warm-up
    best = None
    for rep in range(2):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        c.scan()
        dt = time.perf_counter() - t0
        prof = dict(c.last_scan_profile)
        vals = torch.tensor([dt, prof["search_ms"], prof["allreduce_ms"], prof["unpack_download_ms"]], device=dev,
                            dtype=torch.float64)
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        vals = [float(v) for v in vals.tolist()]
        if best is None or vals[0] < best[0]:
            best = vals
    nonzero, lcc_max = int(np.count_nonzero(c.lcc)), float(c.lcc.max())
    root_best = None
    if world > 1:                                         # the merged grids delivered to rank 0 only (see strong_search)
        c.result_rank = 0
        c.scan()
        for rep in range(2):
            torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            c.scan()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            root_best = float(dt.item()) if root_best is None else min(root_best, float(dt.item()))
        c.result_rank = None
    S = spectrum_bytes(w)
    peak, _ = peak_hbm()
    out = {"workload": w["desc"] + "; %d rotations sharded over %d rank(s)" % (R, world), "rotation_set": rot_desc,
           "rotations": R, "rotations_per_rank": int(c.last_scan_profile["rotations"]), "seconds": best[0],
           "rotations_per_s": R / best[0], "search_ms": best[1], "allreduce_ms": best[2], "unpack_download_ms": best[3],
           "search_frac_per_gpu": int(c.last_scan_profile["rotations"]) / (best[1] / 1e3) * 12 * S / 1e9 / peak,
           "result": {"nonzero": nonzero, "lcc_max": lcc_max}}
    if root_best is not None:
        out.update({"seconds_result_on_rank0": root_best, "rotations_per_s_result_on_rank0": R / root_best})
    del c
    torch.cuda.empty_cache()
    return out


def multi_template_search(dev, rank, world, rot_per_template=2000):
    """BASELINE configs[4] shape: FOUR distinct sub-unit templates against one 192^3 map (plain LCC), `rot_per_template`
    rotations each, through MultiTemplateCorrelator.scan_all(): one plan (FT(map), FT(map^2), work buffers shared),
    one template slot per sub-unit, (template, rotation block) work items dealt over the ranks, ONE MAX all-reduce
    of the [4, V] packed grids.  Set-up outside the timed call; wall clock between barriers, max over ranks."""
    import torch
    import torch.distributed as dist
    from powerfit_b200 import MultiTemplateCorrelator, synth
    n = 192
    case = synth.config5(seed=0)
    subunits = [(case.template, case.mask)] + [
        synth.make_template((n, n, n), 2.0, 8.0, n_res, rg, seed) for n_res, rg, seed in
        ((900, 27.0, 11), (700, 24.0, 12), (500, 21.0, 13))]
    rots = synth.random_rotations(rot_per_template, seed=3)
    m = MultiTemplateCorrelator(case.target, len(subunits), device=dev, laplace=False, shard=True)
    for i, (t, k) in enumerate(subunits):
        m.set_template(i, t, k)
    m.rotations = rots
    m.scan_all()                                         # warm-up
    best = None
    for rep in range(2):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        m.scan_all()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = float(dt.item()) if best is None else min(best, float(dt.item()))
    lcc_max = [float(x.max()) for x in m.lccs]
    root_best = None
    if world > 1:                                        # all four result grids delivered to rank 0 only
        m.result_rank = 0
        m.scan_all()
        for rep in range(2):
            torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            m.scan_all()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            root_best = float(dt.item()) if root_best is None else min(root_best, float(dt.item()))
        m.result_rank = None
    total = len(subunits) * rot_per_template
    S = 8 * n * n * (n // 2 + 1)
    peak, _ = peak_hbm()
    out = {"workload": "192^3 map @8A, batch of %d distinct sub-unit templates, %d rotations each, plain LCC, "
                       "sharded as (template, rotation block) items over %d rank(s)" % (len(subunits), rot_per_template, world),
           "templates": len(subunits), "rotations_total": total, "rotations_this_rank": int(m.last_scan_rotations),
           "seconds": best, "rotations_per_s": total / best,
           "step_frac_per_gpu": total / best / world * 10 * S / 1e9 / peak,
           "lcc_max_per_template": lcc_max}
    if root_best is not None:
        out.update({"seconds_result_on_rank0": root_best, "rotations_per_s_result_on_rank0": total / root_best,
                    "step_frac_per_gpu_result_on_rank0": total / root_best / world * 10 * S / 1e9 / peak})
    del m
    torch.cuda.empty_cache()
    return out


def kernel_split(corr, lib, run_step):
    """Per-kernel-class device time of one extra, untimed step (events around every launch; the scan then keeps
    all kernels on one stream so that the classes do not overlap)."""
    import torch
    from powerfit_b200 import _lib
    kernels = {}
    _lib.check(lib.pfb_profile(corr._plan, 1))
    run_step()
    torch.cuda.synchronize(corr._device)
    _lib.check(lib.pfb_profile(corr._plan, 0))
    cls = 0
    while True:
        kms, kn, name = ctypes.c_double(), ctypes.c_int64(), ctypes.c_char_p()
        if lib.pfb_profile_read(corr._plan, cls, ctypes.byref(kms), ctypes.byref(kn), ctypes.byref(name)) != 0:
            break
        if kn.value:
            kernels[name.value.decode()] = {"ms": kms.value, "launches": kn.value}
        cls += 1
    return kernels


def extra_config(workload, dev, steps=2, warmup=1):
    """Resident-input throughput of another BASELINE config, measured in this process (short run)."""
    import torch
    from powerfit_b200 import CUDACorrelator, _lib
    w = WORKLOADS[workload]
    case = make_inputs(workload)
    rps = default_rps(w)
    rots, rot_desc = search_rotations(workload, rps * (steps + warmup))
    corr = CUDACorrelator(case.target, device=dev, laplace=w["laplace"])
    corr.template, corr.mask, corr.rotations = case.template, case.mask, rots
    nf = 2 if corr._mask_binary else 3
    S = spectrum_bytes(w)
    stream = torch.cuda.current_stream(dev)
    for i in range(warmup):
        corr.scan_device(i * rps, (i + 1) * rps, reset=(i == 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for i in range(warmup, warmup + steps):
        corr.scan_device(i * rps, (i + 1) * rps, reset=False)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    value = rps * steps / (ms / 1e3)
    peak, _ = peak_hbm()
    kern = kernel_split(corr, _lib.load(), lambda: corr.scan_device(0, rps, reset=False))
    out = {"workload": w["desc"], "rotations_per_step": rps, "steps": steps, "warmup": warmup, "rotation_set": rot_desc,
           "shape": list(shape_of(w)), "fused": int(corr.plan_info(6)),
           "value": value, "unit": "rotations/s", "batch": corr.plan_info(4), "mask": "binary" if nf == 2 else "core-weighted",
           "bytes_per_rotation": (2 * nf + 6) * S, "step_frac": value * (2 * nf + 6) * S / 1e9 / peak,
           "us_per_rotation": {k: 1e3 * v["ms"] / rps for k, v in kern.items()}}
    del corr
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--rot-per-step", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip merge_check / strong / the other configs / api_search (kernel experiments)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the one JSON line and nothing else: libraries that write to fd 1 (NCCL prints its version
    # banner there) are sent to stderr, and the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, w, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    from powerfit_b200 import CUDACorrelator, _lib
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = _lib.load()

    # ---------------- N > 1: the merged result is checked before anything is timed
    mcheck = None
    if world > 1 and not args.no_extras:
        mcheck = merge_check(dev, rank, world)
        if not mcheck["ok"]:
            if rank == 0:
                emit({"error": "merge_check failed: sharded search differs from the single-GPU search", "merge_check": mcheck})
            dist.destroy_process_group()
            sys.exit(1)

    case = make_inputs(args.workload)
    n = w["n"]
    V = int(np.prod(shape_of(w)))
    rps = args.rot_per_step or default_rps(w)
    total_steps = args.warmup + args.steps
    rots, rot_desc = search_rotations(args.workload, rps * total_steps * world + 8)

    corr = CUDACorrelator(case.target, device=dev, laplace=w["laplace"], batch=args.batch)
    corr.template = case.template
    corr.mask = case.mask
    corr.rotations = rots
    nf = 2 if corr._mask_binary else 3
    S = spectrum_bytes(w)
    B_rot = (2 * nf + 6) * S
    stream = torch.cuda.current_stream(dev)

    def step(i):
        lo = (i * world + rank) * rps
        best = corr.scan_device(lo, lo + rps, reset=(i == 0))
        if world > 1:
            dist.all_reduce(best, op=dist.ReduceOp.MAX)

    # ---------------- resident-input throughput
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if world > 1:
        dist.barrier()          # rank 0 started the clock sampler: line the ranks up again before timing
    launches0 = corr.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for i in range(args.warmup, total_steps):
        step(i)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    launches = corr.kernel_launches - launches0
    value = world * rps * args.steps / (ms / 1e3)

    # ---------------- end to end through the host-buffer C-ABI call
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    t_laplace = corr._d_target.cpu().numpy()
    h_target, h_lm = pin(t_laplace), pin(corr._lcc_mask)
    h_tmpl, h_mask = pin(corr._template.astype(np.float32)), pin(corr._mask.astype(np.float32))
    h_lcc = torch.empty(V, dtype=torch.float32).pin_memory()
    h_rot = torch.empty(V, dtype=torch.int32).pin_memory()
    e2e_rps = rps
    def e2e_step(i):
        lo = (i * world + rank) * e2e_rps
        sub = np.ascontiguousarray(rots[lo:lo + e2e_rps])
        _lib.check(lib.pfb_search_host(corr._plan, h_target.data_ptr(), h_lm.data_ptr(), h_tmpl.data_ptr(),
                                       h_mask.data_ptr(), float(corr._norm_factor), int(nf == 2),
                                       sub.ctypes.data_as(ctypes.c_void_p), e2e_rps, lo, h_lcc.data_ptr(),
                                       h_rot.data_ptr()))
    for i in range(min(2, args.warmup)):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, total_steps):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_rps * args.steps / float(t_e2e.item())
    h2d = 4 * V * 3 + V + 72 * e2e_rps
    d2h = 8 * V

    # ---------------- the user-level call: CUDACorrelator from host float64 arrays to host lcc/rot grids
    # (plan creation, FP64 preparation on the device, search, download), timed once per preparation mode
    api = {}
    if not args.no_extras:
        sub = rots[:rps]
        for rep in range(3):                     # best of three per mode (the first pays one-time CUDA set-up)
            for mode in ("device", "host"):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                c2 = CUDACorrelator(case.target, device=dev, laplace=w["laplace"], batch=args.batch, prep=mode)
                c2.template, c2.mask, c2.rotations = case.template, case.mask, sub
                c2.scan()
                dt = time.perf_counter() - t0
                if mode not in api or dt < api[mode]["seconds"]:
                    api[mode] = {"seconds": dt, "rotations": int(rps), "rotations_per_s": rps / dt,
                                 "scan_seconds": c2.last_scan_seconds}
                del c2

    # ---------------- per-kernel split (extra, untimed): events around every launch
    kernels = kernel_split(corr, lib, lambda: step(total_steps - 1))
    tot_ms = sum(k["ms"] for k in kernels.values()) or 1.0
    top = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
    peak, peak_src = peak_hbm()
    # algorithmic bytes of each kernel class per rotation (DESIGN.md "roofline")
    alg = {"fused_rotate_fftx": nf * S, "fused_fftyz_mul": (nf + 3) * S, "fused_ifftx_lcc": 3 * S,
           "rotate": nf * S, "fft_x": (2 * nf + 6) * S, "fft_y": (2 * nf + 6) * S, "fft_z": (2 * nf + 6) * S,
           "multiply": (nf + 3) * S, "lcc_best": 3 * S}
    roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
            "traffic_source": None,
            "step_achieved": value / world * B_rot / 1e9, "step_frac": value / world * B_rot / 1e9 / peak,
            "bytes_per_rotation": B_rot}
    batch = corr.plan_info(4)
    if top:
        k = kernels[top]
        ach = alg.get(top, 0) * rps / (k["ms"] / 1e3) / 1e9
        # DRAM bytes of the dominant kernel from the committed ncu --set full capture (per launch of that
        # capture, rescaled to the rotations one launch of this run processes); null if there is no capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[args.workload][top]
            roof["traffic"] = tr["dram_bytes_per_launch"] * min(batch, rps) / tr["rotations_per_launch"]
            roof["traffic_source"] = tr["source"]
        except Exception:
            pass
        roof.update({"kernel": top, "achieved": ach, "frac": ach / peak,
                     "kernel_share_of_step": k["ms"] / tot_ms,
                     "kernel_ms_per_launch": k["ms"] / k["launches"],
                     "kernels": {n_: {"ms_per_step": v["ms"], "launches": v["launches"], "share": v["ms"] / tot_ms,
                                      "us_per_rotation": 1e3 * v["ms"] / rps,
                                      "frac_of_own_bytes": alg.get(n_, 0) * rps / (v["ms"] / 1e3) / 1e9 / peak}
                                 for n_, v in kernels.items()}})
    else:
        roof.update({"achieved": roof["step_achieved"], "frac": roof["step_frac"]})

    del corr
    torch.cuda.empty_cache()

    # ---------------- ONE full search sharded over the ranks (strong scaling), every N
    strong = None
    if not args.no_extras:
        strong = strong_search(dev, rank, world, args.batch)

    # ---------------- the other BASELINE configs, measured in this process (N = 1 only)
    others = {}
    if world == 1 and not args.no_extras:
        for name in ("config1", "config3", "config5", "config4", "cli_96x128x64"):
            if name == args.workload:
                continue
            try:
                others[name] = extra_config(name, dev)
            except Exception as exc:       # never sink the headline
                others[name] = {"error": repr(exc)}

    multi = None
    if not args.no_extras:
        try:
            multi = multi_template_search(dev, rank, world)
        except Exception as exc:           # never sink the headline
            multi = {"error": repr(exc)}

    big = None
    if world > 1 and not args.no_extras:
        try:
            big = sharded_config4(dev, rank, world)
        except Exception as exc:           # never sink the headline
            big = {"error": repr(exc)}

    out = {"metric": "rotations/s (LCC search)", "value": value, "unit": "rotations/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(w, world, rps, batch, rot_desc),
           "e2e": {"value": e2e_val, "unit": "rotations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "call": "pfb_search_host (host buffers in/out, includes FT(map) setup)"},
           "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
           "api_search": dict(api, call="CUDACorrelator(target) -> .template/.mask/.rotations -> .scan() from host "
                                        "float64 arrays, one search of rotations_per_step_per_gpu rotations, "
                                        "per preparation mode (device FP64 kernels / host numpy+scipy)")}
    if mcheck is not None:
        out["merge_check"] = mcheck
    if strong is not None:
        out["strong"] = strong
    if multi is not None:
        out["multi_template"] = multi
    if big is not None:
        out["config4_sharded"] = big
    if others:
        out["configs"] = others
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(case, w["laplace"], rots, n)
            except Exception as exc:       # the baseline must never sink the GPU number
                out["cpu_baseline"] = {"value": None, "error": repr(exc)}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
