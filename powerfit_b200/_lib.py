"""Loader (ctypes) and build recipe for the CUDA library ``libpowerfit_b200.so``.

The library is built in-tree with nvcc for sm_100a only.  There is no CPU or PyTorch
fallback: if the shared object is missing or cannot be loaded, every entry point of
this package raises.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpowerfit_b200.so")
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

# every symbol include/powerfit_b200.h declares
SYMBOLS = ["pfb_version", "pfb_last_error", "pfb_plan_create", "pfb_plan_destroy", "pfb_plan_info",
           "pfb_set_target", "pfb_set_template", "pfb_best_init", "pfb_scan", "pfb_unpack",
           "pfb_merge_best", "pfb_profile", "pfb_profile_read", "pfb_rotate", "pfb_fft3_c2c", "pfb_lcc_take_best", "pfb_search_host",
           "pfb_lcc_max", "pfb_peak_candidates", "pfb_prepare_target", "pfb_prepare_template",
           "pfb_blur_points", "pfb_dilate_points", "pfb_core_indices",
           "pfb_gaussian_filter", "pfb_zoom_linear", "pfb_template_slots", "pfb_select_template", "pfb_pencil_fft"]

_lib = None


class PowerfitB200Error(RuntimeError):
    pass


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libpowerfit_b200.so (nvcc cross-compiles without a GPU)."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(_HERE), "include", "powerfit_b200.h")]
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    # one nvcc per translation unit, side by side (fused.cu alone instantiates ~60 kernels), then one link; the
    # library appears under its final name by rename, so a reader (a loader, a snapshot of the tree) never sees a
    # partial file
    import concurrent.futures
    import shutil
    import tempfile
    objdir = tempfile.mkdtemp(prefix="pfb_build_")
    log = []

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas=-v"] if verbose else []) + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, res

    try:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
            results = list(ex.map(compile_one, srcs))
        for src, obj, res in results:
            if res.returncode != 0:
                raise PowerfitB200Error("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
            log.append(res.stderr)
        tmp = LIB_PATH + ".tmp.%d" % os.getpid()
        res = subprocess.run([nvcc] + NVCC_FLAGS + ["-o", tmp] + [obj for _, obj, _ in results],
                             capture_output=True, text=True)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise PowerfitB200Error("nvcc link failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, LIB_PATH)
    finally:
        shutil.rmtree(objdir, ignore_errors=True)
    if verbose:
        print("\n".join(log))
    return LIB_PATH


def load():
    """Load the CUDA library, or raise (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PowerfitB200Error(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "powerfit_b200 has no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i32, f32 = c.c_void_p, c.c_int, c.c_float
    lib.pfb_version.restype = c.c_char_p
    lib.pfb_last_error.restype = c.c_char_p
    lib.pfb_plan_create.argtypes = [i32, i32, i32, i32, i32, c.POINTER(vp)]
    lib.pfb_plan_destroy.argtypes = [vp]
    lib.pfb_plan_info.argtypes = [vp, i32, c.POINTER(c.c_int64)]
    lib.pfb_set_target.argtypes = [vp, vp, vp, vp]
    lib.pfb_set_template.argtypes = [vp, vp, vp, f32, i32, vp]
    lib.pfb_prepare_target.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.pfb_prepare_template.argtypes = [vp, vp, vp, i32, vp, vp, c.POINTER(c.c_double), c.POINTER(i32), vp]
    lib.pfb_template_slots.argtypes = [vp, i32]
    lib.pfb_select_template.argtypes = [vp, i32]
    lib.pfb_pencil_fft.argtypes = [i32, i32, i32, vp, vp, i32, vp]
    lib.pfb_best_init.argtypes = [vp, vp, vp]
    lib.pfb_scan.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.pfb_unpack.argtypes = [vp, vp, vp, vp, vp]
    lib.pfb_merge_best.argtypes = [vp, vp, vp, vp]
    lib.pfb_profile.argtypes = [vp, i32]
    lib.pfb_profile_read.argtypes = [vp, i32, c.POINTER(c.c_double), c.POINTER(c.c_int64), c.POINTER(c.c_char_p)]
    lib.pfb_rotate.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    lib.pfb_fft3_c2c.argtypes = [vp, vp, i32, vp]
    lib.pfb_lcc_take_best.argtypes = [vp, vp, vp, vp, f32, i32, vp, vp]
    lib.pfb_search_host.argtypes = [vp, vp, vp, vp, vp, f32, i32, vp, i32, i32, vp, vp]
    lib.pfb_lcc_max.argtypes = [vp, c.c_int64, vp, vp, vp]
    lib.pfb_peak_candidates.argtypes = [vp, c.c_int64, f32, i32, vp, vp, vp, vp]
    lib.pfb_blur_points.argtypes = [vp, vp, i32, c.c_double, i32, i32, i32, vp, vp]
    lib.pfb_dilate_points.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    lib.pfb_core_indices.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    lib.pfb_gaussian_filter.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, vp]
    lib.pfb_zoom_linear.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("pfb_version", "pfb_last_error"):
            fn.restype = i32
    _lib = lib
    return lib


def check(rc):
    """Translate a status code: contract violations raise ValueError like the reference's
    correlators do, everything else PowerfitB200Error."""
    if rc == 0:
        return
    msg = load().pfb_last_error().decode("utf-8", "replace")
    if rc == 1:
        raise ValueError(msg)
    raise PowerfitB200Error("powerfit_b200 error %d: %s" % (rc, msg))
