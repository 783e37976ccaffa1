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
rotation.  `cpu_baseline` times the oracle port of the reference CPU path on this box's
cores on a bounded rotation sample.  One JSON line on stdout (rank 0).
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
}


def make_inputs(workload):
    from powerfit_b200 import synth
    w = WORKLOADS[workload]
    if workload == "config1":
        case = synth.config1(seed=0)
    elif workload == "config4":
        case = synth.config4(seed=0)
    elif workload == "config5":
        case = synth.config5(seed=0)
    else:
        case = synth.config2(seed=0, core_weighted=w["cw"])
    return case


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


def cpu_baseline(case, laplace, rotations, per_core=4, cores=None):
    """Oracle port of the reference CPU path (PowerFitter._cpu_scan) on a bounded sample."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=False)
    from oracle import oracle as O
    cores = cores or os.cpu_count() or 1
    n = cores * per_core
    sub = rotations[:n]
    t0 = time.time()
    O.parallel_scan(case.target, case.template, case.mask, sub, laplace=laplace, nproc=cores)
    dt = time.time() - t0
    return {"value": n / dt, "unit": "rotations/s", "cores": cores, "kind": "port",
            "sample": "%d rotations (%d per process) of the same workload, oracle port of CPUCorrelator "
                      "(numpy.fft backend, FP64), %d processes, %.1f s" % (n, per_core, cores, dt)}


def run_reference(args, w, rank, world, emit):
    """--impl reference: the reference's CPU implementation (oracle port; the reference's own
    Python needs its package, which cannot travel) on this box's host cores."""
    if rank != 0:
        return
    from powerfit_b200 import synth
    case = make_inputs(args.workload)
    cores = os.cpu_count() or 1
    per_core = 1 if w["n"] >= 128 else 4
    n = cores * per_core
    rots = synth.random_rotations(n * (args.steps + args.warmup), seed=1)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=False)
    from oracle import oracle as O
    times = []
    for s in range(args.warmup + args.steps):
        sub = rots[s * n:(s + 1) * n]
        t0 = time.time()
        O.parallel_scan(case.target, case.template, case.mask, sub, laplace=w["laplace"], nproc=cores)
        if s >= args.warmup:
            times.append(time.time() - t0)
    T = sum(times)
    val = n * args.steps / T
    sample = "%d rotations per step (%d per process x %d processes), oracle port of CPUCorrelator, numpy.fft FP64" % (
        n, per_core, cores)
    out = {"impl": "reference", "metric": "rotations/s (LCC search)", "value": val, "unit": "rotations/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": w["desc"], "rotations_per_step": n},
           "cpu_baseline": {"value": val, "unit": "rotations/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "rotations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


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
    ap.add_argument("--profile-kernels", action="store_true", help="extra untimed step with per-kernel events")
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
    from powerfit_b200 import CUDACorrelator, _lib, synth
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = _lib.load()

    case = make_inputs(args.workload)
    n = w["n"]
    V = n ** 3
    # one step = one full rotational search of the named angle at 128^3 (7416 rotations); bounded blocks elsewhere
    rps = args.rot_per_step or {64: 2048, 128: w["full_R"], 256: 128, 192: 128}.get(n, 256)
    total_steps = args.warmup + args.steps
    rots = synth.random_rotations(rps * total_steps * world + 8, seed=1)

    corr = CUDACorrelator(case.target, device=dev, laplace=w["laplace"], batch=args.batch)
    corr.template = case.template
    corr.mask = case.mask
    corr.rotations = rots
    nf = 2 if corr._mask_binary else 3
    S = 8 * n * n * (n // 2 + 1)
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
    if rank == 0 or world > 1:
        sub = rots[:rps]
        for rep in range(3):                     # best of three per mode (the first pays one-time CUDA set-up)
            for mode in ("device", "host"):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                c2 = CUDACorrelator(case.target, device=dev, laplace=w["laplace"], batch=args.batch, prep=mode)
                c2.shard = False
                c2.template, c2.mask, c2.rotations = case.template, case.mask, sub
                c2.scan()
                dt = time.perf_counter() - t0
                if mode not in api or dt < api[mode]["seconds"]:
                    api[mode] = {"seconds": dt, "rotations": int(rps), "rotations_per_s": rps / dt,
                                 "scan_seconds": c2.last_scan_seconds}
                del c2

    # ---------------- per-kernel split (extra, untimed): events around every launch
    kernels = {}
    _lib.check(lib.pfb_profile(corr._plan, 1))
    step(total_steps - 1)
    torch.cuda.synchronize(dev)
    _lib.check(lib.pfb_profile(corr._plan, 0))
    cls = 0
    while True:
        kms, kn, name = ctypes.c_double(), ctypes.c_int64(), ctypes.c_char_p()
        if lib.pfb_profile_read(corr._plan, cls, ctypes.byref(kms), ctypes.byref(kn), ctypes.byref(name)) != 0:
            break
        if kn.value:
            kernels[name.value.decode()] = {"ms": kms.value, "launches": kn.value}
        cls += 1
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
    if top:
        k = kernels[top]
        ach = alg.get(top, 0) * rps / (k["ms"] / 1e3) / 1e9
        # DRAM bytes of the dominant kernel from the committed ncu --set full capture (per launch of that
        # capture, rescaled to the rotations one launch of this run processes); null if there is no capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[args.workload][top]
            roof["traffic"] = tr["dram_bytes_per_launch"] * min(corr.plan_info(4), rps) / tr["rotations_per_launch"]
            roof["traffic_source"] = tr["source"]
        except Exception:
            pass
        roof.update({"kernel": top, "achieved": ach, "frac": ach / peak,
                     "kernel_share_of_step": k["ms"] / tot_ms,
                     "kernel_ms_per_launch": k["ms"] / k["launches"],
                     "kernels": {n_: {"ms_per_step": v["ms"], "launches": v["launches"], "share": v["ms"] / tot_ms}
                                 for n_, v in kernels.items()}})
    else:
        roof.update({"achieved": roof["step_achieved"], "frac": roof["step_frac"]})

    out = {"metric": "rotations/s (LCC search)", "value": value, "unit": "rotations/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": w["desc"], "shape": [n, n, n], "rotations_per_step_per_gpu": rps,
                      "batch": corr.plan_info(4), "mask": "binary" if nf == 2 else "core-weighted",
                      "laplace": w["laplace"], "parallelism": "rotation shards x%d + packed MAX all-reduce" % world,
                      "l2": "working set per batch (%.0f MB) exceeds the 126 MB L2; no flush needed"
                            % (corr.plan_info(4) / 2 * (nf + 3) * 8 * V / 1e6)},
           "e2e": {"value": e2e_val, "unit": "rotations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "call": "pfb_search_host (host buffers in/out, includes FT(map) setup)"},
           "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
           "api_search": dict(api, call="CUDACorrelator(target) -> .template/.mask/.rotations -> .scan() from host "
                                        "float64 arrays, one search of rotations_per_step_per_gpu rotations, "
                                        "per preparation mode (device FP64 kernels / host numpy+scipy)")}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(case, w["laplace"], rots, per_core=2 if n >= 128 else 8)
            except Exception as exc:       # the baseline must never sink the GPU number
                out["cpu_baseline"] = {"value": None, "error": repr(exc)}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
