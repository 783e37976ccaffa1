"""``CUDACorrelator`` -- the B200 drop-in for the reference's ``GPUCorrelator``.

Same duck-typed contract as ``BaseCorrelator``/``GPUCorrelator``
(/root/reference/src/powerfit_em/powerfitter.py:166-255, 396-555): constructor takes
the target array (+ a device handle + ``laplace``), then ``.template``, ``.mask``,
``.rotations`` are set in that order, ``.scan()`` runs the search and ``.lcc``
(float32) / ``.rot`` (int32) hold the results.  The same ``ValueError`` messages are
raised for the same contract violations.

One-time preparation (target normalisation, lcc_mask, Laplace filter, template
z-scoring) follows the reference formulas in FP64 and is then cast to FP32, like
``GPUCorrelator`` does (powerfitter.py:414, 471-474).  By default it runs on the device
(``pfb_prepare_target`` / ``pfb_prepare_template``: the FP64 arrays are uploaded as they
are); ``prep="host"`` keeps the numpy/scipy formulation, which the tests use as the
cross-check of the device kernels.  Everything per-rotation runs in
the hand-written sm_100a kernels behind the C ABI (``include/powerfit_b200.h``);
PyTorch only provides device buffers, the stream and -- when a process group is
initialised -- the single MAX all-reduce that merges the rotation shards.
"""
from __future__ import annotations

import ctypes
import os
from sys import stdout
from time import time

import numpy as np

from . import _lib


def _laplace_wrap(a):
    """powerfitter.py:212-215 (scipy.ndimage.laplace, mode='wrap')."""
    from scipy.ndimage import laplace
    return laplace(a, mode="wrap")


def shard_bounds(nrot, world, rank):
    """Contiguous rotation blocks of nrot//world, last rank takes the remainder
    (powerfitter.py:95-108)."""
    per = nrot // world
    lo = rank * per
    hi = nrot if rank == world - 1 else lo + per
    return lo, hi


FUSED_AXES = (32, 64, 96, 128)        # any mix of these per axis (csrc/fused.cu)
FUSED_CUBES = (192, 256)              # cubes only (csrc/fused_cls.cu)


def fused_shape(shape):
    """Smallest grid with a fused pipeline that holds `shape`: every axis rounded up to 32, 64, 96 or 128 voxels,
    a 192^3 or 256^3 cube when an axis is longer than 128, None above 256."""
    if max(shape) <= FUSED_AXES[-1]:
        return tuple(min(n for n in FUSED_AXES if n >= s) for s in shape)
    for n in FUSED_CUBES:
        if max(shape) <= n:
            return (n, n, n)
    return None


def fused_cube(shape):
    """Edge of the smallest CUBIC grid with a fused pipeline that holds `shape`, or None."""
    for n in (64, 128, 192, 256):
        if max(shape) <= n:
            return n
    return None


def _as_shape(n):
    return (n, n, n) if np.isscalar(n) else tuple(int(v) for v in n)


def pad_target(a, n):
    """Zero-pad a map at the high end of every axis to the shape n (an int means n^3) -- what the reference CLI's
    `extend` does (volume.py:102-118), just further."""
    out = np.zeros(_as_shape(n), dtype=a.dtype)
    out[:a.shape[0], :a.shape[1], :a.shape[2]] = a
    return out


def pad_wrapped(a, n):
    """Zero-pad a template / mask that is centred on voxel 0 with wrap-around: non-negative offsets stay at
    the low end of every axis, negative offsets move to the high end of the grid of shape n (an int means n^3)."""
    shape = _as_shape(n)
    out = np.zeros(shape, dtype=a.dtype)
    cuts = [s // 2 + 1 for s in a.shape]
    for sz in (0, 1):
        for sy in (0, 1):
            for sx in (0, 1):
                src, dst = [], []
                for ax, side in enumerate((sz, sy, sx)):
                    h, length = cuts[ax], a.shape[ax]
                    if side == 0:
                        src.append(slice(0, h)); dst.append(slice(0, h))
                    else:
                        src.append(slice(h, length)); dst.append(slice(shape[ax] - (length - h), shape[ax]))
                out[tuple(dst)] = a[tuple(src)]
    return out


def _resolve_device(device):
    import torch
    if not torch.cuda.is_available():
        raise _lib.PowerfitB200Error("CUDACorrelator needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        if "LOCAL_RANK" in os.environ:
            return torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        return torch.device("cuda", torch.cuda.current_device())
    if isinstance(device, int):
        return torch.device("cuda", device)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ValueError("device must be a CUDA device")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class CUDACorrelator(object):
    """B200 implementation of the local cross-correlation search."""

    def __init__(self, target, device=None, laplace=False, batch=0, prep="device", pad=False, shard=False,
                 group=None, result_rank=None):
        """``shard=True`` (opt-in): split the rotation list over the ranks of the torch.distributed process
        group ``group`` (default: the world group) and merge with one MAX all-reduce, see ``scan``.
        ``result_rank=r`` (with ``shard=True``): only rank ``r`` of the group receives the merged grids -- a MAX
        reduce to that rank instead of the all-reduce, and the other ranks skip the unpacking and the download
        (their ``.lcc`` / ``.rot`` are None).  This is what the reference's multi-process search does: the worker
        processes hand their partial grids to the parent (powerfitter.py:135-163), nobody else sees the result.

        ``pad=True`` (opt-in, not reference behaviour): zero-pad the map to the next grid that has a fused
        pipeline (every axis up to 32, 64, 96 or 128 voxels; 192^3 / 256^3 beyond) and crop the results back.
        Grids whose axes already are such lengths take the fused pipeline without it.  The result is exactly
        the search on the padded map -- what the reference computes when its own `extend` step
        (powerfit.py:230-233) is given that size -- and about ten times faster than the any-shape pipeline;
        near the box faces it differs from the unpadded search, whose template wraps around."""
        import torch
        self._torch = torch
        self._libh = _lib.load()
        target = np.asarray(target, dtype=np.float64)
        if target.ndim != 3:
            raise ValueError("target must be a 3-D array")
        self._crop = None
        if pad:
            n = fused_shape(target.shape)
            if n is not None and target.shape != n:
                self._crop = target.shape
                target = pad_target(target, n)
        if prep not in ("device", "host"):
            raise ValueError("prep must be 'device' or 'host'")
        self._prep = prep
        self._target_in = target                       # normalised lazily (property _target)
        self._shape = target.shape
        self._target_h = None
        self._lcc_mask_h = None
        self._rotations = None
        self._template_h = None
        self._template_set = False
        self._mask = None
        self._mask_binary = None
        self._laplace = laplace
        self._rmax = min(target.shape) // 2            # powerfitter.py:176
        self._lcc = None
        self._rot = None
        self.progress = False
        self.shard = bool(shard)    # split rotations over torch.distributed ranks (every rank must call scan())
        self.result_rank = None if result_rank is None else int(result_rank)
        self.group = group
        self.last_scan_seconds = None
        self.last_scan_profile = None

        self._device = _resolve_device(device)
        self._plan = ctypes.c_void_p()
        nz, ny, nx = target.shape
        _lib.check(self._libh.pfb_plan_create(nz, ny, nx, int(batch), self._device.index,
                                              ctypes.byref(self._plan)))
        with torch.cuda.device(self._device):
            self._best = torch.empty(target.size, dtype=torch.int64, device=self._device)
            if prep == "device":
                # BaseCorrelator.__init__ + GPUCorrelator.__init__ (powerfitter.py:169-180, 410-414) in FP64 kernels
                d64 = torch.from_numpy(np.ascontiguousarray(target)).to(self._device)
                self._d_target = torch.empty(target.shape, dtype=torch.float32, device=self._device)
                self._d_lcc_mask = torch.empty(target.shape, dtype=torch.uint8, device=self._device)
                _lib.check(self._libh.pfb_prepare_target(self._plan, d64.data_ptr(), int(bool(laplace)),
                                                         self._d_target.data_ptr(), self._d_lcc_mask.data_ptr(),
                                                         self._stream()))
                torch.cuda.current_stream(self._device).synchronize()
            else:
                t = _laplace_wrap(self._target) if laplace else self._target      # powerfitter.py:410-412
                self._d_target = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32)).to(self._device)
                self._d_lcc_mask = torch.from_numpy(np.ascontiguousarray(self._lcc_mask)).to(self._device)
                _lib.check(self._libh.pfb_set_target(self._plan, self._d_target.data_ptr(),
                                                     self._d_lcc_mask.data_ptr(), self._stream()))

    # host views of the prepared inputs, materialised on first use
    @property
    def _target(self):                                 # powerfitter.py:170
        if self._target_h is None:
            self._target_h = self._target_in / self._target_in.max()
        return self._target_h

    @property
    def _lcc_mask(self):                               # powerfitter.py:178-180
        if self._lcc_mask_h is None:
            if self._prep == "device":
                self._lcc_mask_h = self._d_lcc_mask.cpu().numpy()
            else:
                self._lcc_mask_h = (self._target > self._target.max() * 0.05).astype(np.uint8)
        return self._lcc_mask_h

    @property
    def _template(self):
        if self._template_h is None and self._template_set and self._prep == "device" and self._mask is not None:
            self._template_h = self._d_template.cpu().numpy().astype(np.float64)
        return self._template_h

    @_template.setter
    def _template(self, value):
        self._template_h = value

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                self._libh.pfb_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(self._torch.cuda.current_stream(self._device).cuda_stream)

    def plan_info(self, what):
        v = ctypes.c_int64()
        _lib.check(self._libh.pfb_plan_info(self._plan, what, ctypes.byref(v)))
        return v.value

    @property
    def kernel_launches(self):
        return self.plan_info(7)

    # ------------------------------------------------------------------ contract
    @property
    def target(self):
        return self._target

    @property
    def template(self):
        return self._template

    @template.setter
    def template(self, template):                      # powerfitter.py:236-243
        template = np.asarray(template)
        if self._crop is not None and template.shape == self._crop:
            template = pad_wrapped(template, self._shape)
        if template.shape != self._shape:
            raise ValueError("Shape of template does not match the target.")
        self._mask = None
        self._template_set = True
        self._template = np.array(template, dtype=np.float64)

    @property
    def mask(self):
        return self._mask

    @mask.setter
    def mask(self, mask):                              # powerfitter.py:190-210, 466-474
        if not self._template_set:
            raise ValueError("First set the template.")
        mask = np.asarray(mask)
        if self._crop is not None and mask.shape == self._crop:
            mask = pad_wrapped(mask, self._shape)
        if self._shape != mask.shape:
            raise ValueError("Shape of the mask is different from target.")
        torch = self._torch
        if self._prep == "device":
            m64 = np.array(mask, dtype=np.float64)
            with torch.cuda.device(self._device):
                d_t64 = torch.from_numpy(np.ascontiguousarray(self._template)).to(self._device)
                d_m64 = torch.from_numpy(m64).to(self._device)
                self._d_template = torch.empty(self._shape, dtype=torch.float32, device=self._device)
                self._d_mask = torch.empty(self._shape, dtype=torch.float32, device=self._device)
                nf, binary = ctypes.c_double(), ctypes.c_int()
                _lib.check(self._libh.pfb_prepare_template(
                    self._plan, d_t64.data_ptr(), d_m64.data_ptr(), int(bool(self._laplace)),
                    self._d_template.data_ptr(), self._d_mask.data_ptr(), ctypes.byref(nf), ctypes.byref(binary),
                    self._stream()))
                torch.cuda.current_stream(self._device).synchronize()
            self._norm_factor = int(nf.value)
            self._mask_binary = bool(binary.value)
            self._mask = m64
            self._template_h = None                    # prepared template: fetched from the device on demand
            return
        ind = mask != 0
        self._norm_factor = ind.sum()
        if self._norm_factor == 0:
            raise ValueError("Zero-filled mask is not allowed.")
        self._mask = np.array(mask, dtype=np.float64)
        t = self._template
        if self._laplace:
            t = _laplace_wrap(t)
        t *= self._mask
        t[ind] -= t[ind].mean()                        # powerfitter.py:217-220
        t[ind] /= t[ind].std()
        t *= self._mask
        self._template_h = t
        binary = bool(np.all(self._mask[ind] == 1.0))
        self._mask_binary = binary
        with torch.cuda.device(self._device):
            self._d_template = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32)).to(self._device)
            self._d_mask = torch.from_numpy(np.ascontiguousarray(self._mask, dtype=np.float32)).to(self._device)
            _lib.check(self._libh.pfb_set_template(self._plan, self._d_template.data_ptr(), self._d_mask.data_ptr(),
                                                   float(self._norm_factor), int(binary), self._stream()))
            torch.cuda.current_stream(self._device).synchronize()

    @property
    def rotations(self):
        return self._rotations

    @rotations.setter
    def rotations(self, rotations):                    # powerfitter.py:226-230
        self._rotations = np.ascontiguousarray(
            np.asarray(rotations, dtype=np.float64).reshape(-1, 3, 3))

    @property
    def lcc(self):
        return self._lcc

    @property
    def rot(self):
        return self._rot

    # ------------------------------------------------------------------ search
    def scan_device(self, lo=None, hi=None, reset=True):
        """Run rotations [lo, hi) (global indices) into the device-resident packed best
        grid; no host transfer except the rotation matrices.  Returns the int64 tensor."""
        if not self._template_set or self._mask is None or self._rotations is None:
            raise ValueError("First set the template, mask, and rotations.")
        nrot = self._rotations.shape[0]
        lo = 0 if lo is None else lo
        hi = nrot if hi is None else hi
        torch = self._torch
        with torch.cuda.device(self._device):
            s = self._stream()
            if reset:
                _lib.check(self._libh.pfb_best_init(self._plan, self._best.data_ptr(), s))
            if self.progress:
                step = max(2 * self.plan_info(4), (hi - lo) // 200)
                time0 = time()
                for a in range(lo, hi, step):
                    b = min(hi, a + step)
                    self._scan_block(a, b, s)
                    torch.cuda.current_stream(self._device).synchronize()
                    self._print_progress(b - lo - 1, hi - lo, time0)
                stdout.write("\n")
            else:
                self._scan_block(lo, hi, s)
        return self._best

    def _scan_block(self, a, b, s):
        sub = self._rotations[a:b]
        _lib.check(self._libh.pfb_scan(self._plan, sub.ctypes.data_as(ctypes.c_void_p), b - a, a,
                                       self._best.data_ptr(), s))

    def scan(self):
        """GPUCorrelator.scan (powerfitter.py:513-538).  With ``self.shard`` set and an initialised
        torch.distributed process group each rank searches its contiguous block of the rotation list
        (powerfitter.py:95-108) and the packed best grids are merged by a single integer MAX all-reduce
        (the order-preserving key reproduces the reference's merge, powerfitter.py:146-163); every rank
        ends with the full result (or, with ``result_rank`` set, only that rank: a MAX reduce, no unpacking and
        no download elsewhere -- the reference's parent-process semantics).  EVERY rank of the group must call scan() with the same target,
        template, mask and rotations -- sharding is therefore opt-in (``shard=True`` in the constructor
        or ``PowerFitter(..., shard=True)``); without it a rank searches the whole list on its own.
        ``last_scan_profile`` holds the device-timed split of the call (search / all-reduce / unpack+download)."""
        torch = self._torch
        t0 = time()
        nrot = 0 if self._rotations is None else self._rotations.shape[0]
        world, rank = 1, 0
        dist = torch.distributed
        if self.shard and dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = shard_bounds(nrot, world, rank)
        stream = torch.cuda.current_stream(self._device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.device(self._device):
            ev[0].record(stream)
            best = self.scan_device(lo, hi)
            ev[1].record(stream)
            mine = True                                # does this rank receive the merged grids?
            if world > 1:
                if self.result_rank is None:
                    dist.all_reduce(best, op=dist.ReduceOp.MAX, group=self.group)
                else:
                    dst = self.result_rank if self.group is None else dist.get_global_rank(self.group, self.result_rank)
                    dist.reduce(best, dst=dst, op=dist.ReduceOp.MAX, group=self.group)
                    mine = rank == self.result_rank
            ev[2].record(stream)
            if not mine:
                ev[3].record(stream)
                ev[3].synchronize()
                self._lcc = self._rot = None
                self.last_scan_seconds = time() - t0
                self.last_scan_profile = {"rotations": int(hi - lo), "world": int(world),
                                          "search_ms": ev[0].elapsed_time(ev[1]),
                                          "allreduce_ms": ev[1].elapsed_time(ev[2]), "unpack_download_ms": 0.0}
                return
            # both grids in one device buffer -> ONE DMA transfer into page-locked host memory (powerfitter.py:536-537)
            V = int(np.prod(self._shape))
            out = torch.empty(2 * V, dtype=torch.int32, device=self._device)
            lcc, rot = out[:V].view(torch.float32), out[V:]
            _lib.check(self._libh.pfb_unpack(self._plan, best.data_ptr(), lcc.data_ptr(), rot.data_ptr(),
                                             self._stream()))
            host, h = self._result_buffer(2 * V)
            host.copy_(out, non_blocking=True)
            ev[3].record(stream)
            ev[3].synchronize()
            self._lcc = h[:V].view(np.float32).reshape(self._shape)
            self._rot = h[V:].reshape(self._shape)
            del h
            if self._crop is not None:
                nz, ny, nx = self._crop
                self._lcc = np.ascontiguousarray(self._lcc[:nz, :ny, :nx])
                self._rot = np.ascontiguousarray(self._rot[:nz, :ny, :nx])
        self.last_scan_seconds = time() - t0
        self.last_scan_profile = {"rotations": int(hi - lo), "world": int(world),
                                  "search_ms": ev[0].elapsed_time(ev[1]), "allreduce_ms": ev[1].elapsed_time(ev[2]),
                                  "unpack_download_ms": ev[2].elapsed_time(ev[3])}

    def _result_buffer(self, count):
        """Page-locked int32 host buffer for the result grids, as (tensor, numpy base array).  `.lcc` / `.rot` are
        views of the base array, so the buffer of the previous scan is reused only when nothing outside this object
        still refers to it (the caller may have kept the earlier arrays, or views of them, the way it can keep the
        fresh arrays the reference returns); otherwise a new buffer is allocated."""
        import sys
        torch = self._torch
        prev = getattr(self, "_host_result", None)
        if prev is not None and prev[0].numel() == count:
            self._lcc = self._rot = None
            if sys.getrefcount(prev[1]) <= 2:          # the tuple's reference + the call argument: no view is alive
                return prev
        t = torch.empty(count, dtype=torch.int32).pin_memory()
        self._host_result = (t, t.numpy())
        return self._host_result

    @staticmethod
    def _print_progress(n, nrot, time0):               # powerfitter.py:540-547
        p_done = (n + 1) / float(nrot) * 100
        now = time()
        eta = ((now - time0) / p_done) * (100 - p_done)
        total = (now - time0) / p_done * (100)
        stdout.write("{:7.2%} {:.0f}s {:.0f}s       \r".format(n / float(nrot), eta, total))
        stdout.flush()

    # ------------------------------------------------------------------ operator-level access
    def rotate(self, grid, rotmats, nearest=False):
        """Device twin of _extensions.rotate_grid3d for R rotations -> (R, nz, ny, nx) float32."""
        torch = self._torch
        rotmats = np.ascontiguousarray(np.asarray(rotmats, dtype=np.float64).reshape(-1, 3, 3))
        R = rotmats.shape[0]
        with torch.cuda.device(self._device):
            g = torch.from_numpy(np.ascontiguousarray(grid, dtype=np.float32)).to(self._device)
            out = torch.empty((R,) + tuple(self._shape), dtype=torch.float32, device=self._device)
            _lib.check(self._libh.pfb_rotate(self._plan, g.data_ptr(), rotmats.ctypes.data_as(ctypes.c_void_p),
                                             R, int(bool(nearest)), out.data_ptr(), self._stream()))
            return out.cpu().numpy()

    def fft3(self, vols):
        """In-place-style 3-D complex DFT, kernel exp(+2 pi i k r / n), un-normalised."""
        torch = self._torch
        v = np.ascontiguousarray(vols, dtype=np.complex64)
        nvol = v.size // int(np.prod(self._shape))
        with torch.cuda.device(self._device):
            d = torch.from_numpy(v.view(np.float32)).to(self._device)
            _lib.check(self._libh.pfb_fft3_c2c(self._plan, d.data_ptr(), nvol, self._stream()))
            return d.cpu().numpy().view(np.complex64).reshape(v.shape)

    def lcc_take_best(self, gcc, ave, ave2, norm_factor, rot_index, best=None):
        """Device twin of CLKernels.calc_lcc_and_take_best on host arrays; returns
        (lcc, rot, best_tensor)."""
        torch = self._torch
        with torch.cuda.device(self._device):
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self._device)
            g, a1, a2 = up(gcc), up(ave), up(ave2)
            if best is None:
                best = torch.empty(int(np.prod(self._shape)), dtype=torch.int64, device=self._device)
                _lib.check(self._libh.pfb_best_init(self._plan, best.data_ptr(), self._stream()))
            _lib.check(self._libh.pfb_lcc_take_best(self._plan, g.data_ptr(), a1.data_ptr(), a2.data_ptr(),
                                                    float(norm_factor), int(rot_index), best.data_ptr(),
                                                    self._stream()))
            lcc = torch.empty(self._shape, dtype=torch.float32, device=self._device)
            rot = torch.empty(self._shape, dtype=torch.int32, device=self._device)
            _lib.check(self._libh.pfb_unpack(self._plan, best.data_ptr(), lcc.data_ptr(), rot.data_ptr(),
                                             self._stream()))
            return lcc.cpu().numpy(), rot.cpu().numpy(), best


def template_work_items(ntemplates, nrot, world):
    """(template, lo, hi) work items of a multi-template search for `world` ranks: every template's rotation list
    is cut into K = world / gcd(T, world) contiguous blocks (powerfitter.py:95-108 per template), so that the
    T K items divide evenly over the ranks; item i belongs to rank i % world."""
    from math import gcd
    k = world // gcd(ntemplates, world)
    items = []
    for t in range(ntemplates):
        for b in range(k):
            lo, hi = shard_bounds(nrot, k, b)
            items.append((t, lo, hi))
    return items


class MultiTemplateCorrelator(CUDACorrelator):
    """Several templates against ONE map (BASELINE configs[4]: fitting a batch of sub-units).

    The reference runs one ``PowerFitter`` -- one correlator, one FT(map), FT(map^2) -- per template
    (powerfit.py:245-282).  Here the plan keeps one template slot per sub-unit (``pfb_template_slots`` /
    ``pfb_select_template``) and shares the map spectra, lcc_mask and work buffers between them; every slot is
    filled through the same ``.template`` / ``.mask`` setters as a single-template correlator (same preparation,
    same errors) and gives exactly the result a fresh ``CUDACorrelator`` gives for that template.

    ``scan_all()`` searches every template over ``.rotations``.  With ``shard=True`` under torch.distributed the
    (template, rotation block) work items are dealt over the ranks and ALL templates' packed best grids are
    merged by one MAX all-reduce of the [T, V] int64 tensor (``result_rank=r``: a MAX reduce to rank r, which alone
    unpacks and downloads the grids)."""

    _SLOT_ATTRS = ("_template_h", "_template_set", "_mask", "_mask_binary", "_norm_factor", "_d_template", "_d_mask")

    def __init__(self, target, ntemplates, **kw):
        super().__init__(target, **kw)
        if ntemplates < 1:
            raise ValueError("ntemplates must be positive")
        _lib.check(self._libh.pfb_template_slots(self._plan, int(ntemplates)))
        self.ntemplates = int(ntemplates)
        self._slot = 0
        self._slot_state = [None] * self.ntemplates
        self.lccs = [None] * self.ntemplates
        self.rots = [None] * self.ntemplates

    def select(self, slot):
        if not 0 <= slot < self.ntemplates:
            raise ValueError("no such template slot")
        if slot == self._slot:
            return
        self._slot_state[self._slot] = {k: getattr(self, k, None) for k in self._SLOT_ATTRS}
        _lib.check(self._libh.pfb_select_template(self._plan, int(slot)))
        state = self._slot_state[slot] or {"_template_h": None, "_template_set": False, "_mask": None,
                                           "_mask_binary": None, "_norm_factor": None, "_d_template": None,
                                           "_d_mask": None}
        for k, v in state.items():
            setattr(self, k, v)
        self._slot = slot

    def set_template(self, slot, template, mask):
        self.select(slot)
        self.template = template
        self.mask = mask

    def scan_all(self):
        torch = self._torch
        if self._rotations is None:
            raise ValueError("First set the template, mask, and rotations.")
        t0 = time()
        nrot, T = self._rotations.shape[0], self.ntemplates
        world, rank = 1, 0
        dist = torch.distributed
        if self.shard and dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        V = int(np.prod(self._shape))
        with torch.cuda.device(self._device):
            best = torch.empty((T, V), dtype=torch.int64, device=self._device)
            for t in range(T):
                _lib.check(self._libh.pfb_best_init(self._plan, best[t].data_ptr(), self._stream()))
            mine = template_work_items(T, nrot, world)[rank::world]
            done = 0
            for t, lo, hi in mine:
                self.select(t)
                if not self._template_set or self._mask is None:
                    raise ValueError("First set the template, mask, and rotations.")
                sub = self._rotations[lo:hi]
                _lib.check(self._libh.pfb_scan(self._plan, sub.ctypes.data_as(ctypes.c_void_p), hi - lo, lo,
                                               best[t].data_ptr(), self._stream()))
                done += hi - lo
            if world > 1:
                if self.result_rank is None:
                    dist.all_reduce(best, op=dist.ReduceOp.MAX, group=self.group)
                else:                                  # results on one rank only (see CUDACorrelator)
                    dst = self.result_rank if self.group is None else dist.get_global_rank(self.group, self.result_rank)
                    dist.reduce(best, dst=dst, op=dist.ReduceOp.MAX, group=self.group)
                    if rank != self.result_rank:
                        torch.cuda.current_stream(self._device).synchronize()
                        self.lccs, self.rots = [None] * T, [None] * T
                        self._lcc = self._rot = None
                        self.last_scan_seconds = time() - t0
                        self.last_scan_rotations = done
                        return
            # all grids in one device buffer -> one DMA transfer into page-locked host memory
            out = torch.empty(2 * T * V, dtype=torch.int32, device=self._device)
            d_lcc, d_rot = out[:T * V].view(torch.float32).view(T, V), out[T * V:].view(T, V)
            for t in range(T):
                _lib.check(self._libh.pfb_unpack(self._plan, best[t].data_ptr(), d_lcc[t].data_ptr(),
                                                 d_rot[t].data_ptr(), self._stream()))
            self.lccs = [None] * T
            self.rots = [None] * T
            host, h = self._result_buffer(2 * T * V)
            host.copy_(out, non_blocking=True)
            torch.cuda.current_stream(self._device).synchronize()
            lcc = h[:T * V].view(np.float32).reshape((T,) + tuple(self._shape))
            rot = h[T * V:].reshape((T,) + tuple(self._shape))
            del h
        for t in range(T):
            self.lccs[t], self.rots[t] = lcc[t], rot[t]
            if self._crop is not None:
                nz, ny, nx = self._crop
                self.lccs[t] = np.ascontiguousarray(lcc[t][:nz, :ny, :nx])
                self.rots[t] = np.ascontiguousarray(rot[t][:nz, :ny, :nx])
        self.last_scan_seconds = time() - t0
        self.last_scan_rotations = done
