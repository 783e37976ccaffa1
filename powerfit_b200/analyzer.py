"""``Analyzer`` -- solution extraction after a search.

Same constructor, properties, ``solutions`` rows and ``solutions.out`` format as
/root/reference/src/powerfit_em/analyzer.py:4-113, so the reference CLI
(powerfit.py:285-305) can use it unchanged.  What is different is where the work happens:
the reference labels ``corr >= cutoff`` with ``scipy.ndimage.label`` for each of ``steps``
cutoffs -- ``steps`` passes over the whole grid.  A voxel below the lowest cutoff is never
part of a feature, so here the device finds the grid maximum and compacts the voxels at or
above the lowest cutoff into a short (index, value) list (csrc/peaks.cu, `pfb_lcc_max`,
`pfb_peak_candidates`); the 6-connected components of that list at every cutoff, and the
position of the maximum of each, are then found on the list (a few thousand entries)
instead of the grid.  The positions are the ones ``label`` + ``maximum_position`` return
(analyzer.py:90-94; tests pin this against the reference's own output).

There is no CPU fallback for the grid-sized part: a missing library or GPU raises.
"""
from __future__ import annotations

import numpy as np


def watershed_cutoffs(max_cc, steps):
    """The ``steps`` descending cutoffs of analyzer.py:84-90, in the arithmetic (dtype) the
    reference uses: everything stays in the dtype of ``corr.max()``."""
    max_cc = np.asarray(max_cc)[()]
    min_cc = 0.5 * max_cc
    stepsize = (max_cc - min_cc) / steps
    cutoff = max_cc
    out = []
    for _ in range(steps):
        cutoff = cutoff - stepsize
        out.append(cutoff)
    return out


def sparse_feature_maxima(idx, val, shape, cutoffs):
    """Positions (linear C-order indices) of the maximum of every 6-connected feature of
    ``{corr >= cutoff}``, for every cutoff, given only the voxels at or above the lowest
    cutoff: ``idx`` (linear indices, any order) and ``val`` (their values).

    Equivalent to ``scipy.ndimage.label`` (default cross-shaped structure, no periodic
    wrap) followed by ``maximum_position`` per label (analyzer.py:92-93); among exactly equal
    maxima of one feature the lowest linear index is returned."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    idx = np.asarray(idx, dtype=np.int64)
    val = np.asarray(val)
    order = np.argsort(idx, kind="stable")
    idx, val = idx[order], val[order]
    n = idx.size
    if n == 0:
        return set()
    nz, ny, nx = shape
    x = idx % nx
    y = (idx // nx) % ny
    z = idx // (nx * ny)
    # edges to the +x, +y, +z neighbour when that neighbour is in the list as well
    src, dst = [], []
    for step, ok in ((1, x < nx - 1), (nx, y < ny - 1), (nx * ny, z < nz - 1)):
        want = idx + step
        pos = np.searchsorted(idx, want)
        pos[pos >= n] = n - 1
        hit = ok & (idx[pos] == want)
        src.append(np.nonzero(hit)[0])
        dst.append(pos[hit])
    src = np.concatenate(src)
    dst = np.concatenate(dst)
    found = set()
    for cutoff in cutoffs:
        live = val >= cutoff
        if not live.any():
            continue
        e = live[src] & live[dst]
        graph = coo_matrix((np.ones(int(e.sum()), dtype=np.int8), (src[e], dst[e])), shape=(n, n))
        _, lab = connected_components(graph, directed=False)
        members = np.nonzero(live)[0]
        # per label: highest value, ties -> lowest linear index (members are index-sorted)
        key = np.lexsort((members, -val[members].astype(np.float64), lab[members]))
        srt = members[key]
        first = np.ones(srt.size, dtype=bool)
        first[1:] = lab[srt][1:] != lab[srt][:-1]
        found.update(int(i) for i in idx[srt[first]])
    return found


class Analyzer(object):

    def __init__(self, corr, rotmat, rotmat_ind, steps=5, voxelspacing=1,
                 origin=(0, 0, 0), z_sigma=1, device=None):
        self._corr = corr
        self._rotmat = rotmat
        self._rotmat_ind = rotmat_ind
        self._voxelspacing = voxelspacing
        self._origin = origin
        self._z_sigma = z_sigma
        self._device = device
        self.last_candidates = 0
        self.steps = steps
        self._solutions = None

    @property
    def corr(self):
        return self._corr

    @property
    def steps(self):
        return self._steps

    @steps.setter
    def steps(self, steps):
        self._steps = steps
        self._watershed()
        self._solutions = None

    @property
    def voxelspacing(self):
        return self._voxelspacing

    @voxelspacing.setter
    def voxelspacing(self, voxelspacing):
        self._solutions = None
        self._voxelspacing = voxelspacing

    @property
    def origin(self):
        return self._origin

    @origin.setter
    def origin(self, origin):
        self._solutions = None
        self._origin = origin

    @property
    def solutions(self):
        if self._solutions is None:
            self._generate_solutions()
        return self._solutions

    # ------------------------------------------------------------------ device part
    def _device_candidates(self):
        """(max, idx, val) of the voxels at or above the lowest cutoff -- two kernels over the
        grid; ``corr`` may be a float32 numpy array (uploaded once) or a CUDA torch tensor."""
        import torch
        from . import _lib
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.PowerfitB200Error("powerfit_b200.Analyzer needs a CUDA device (no CPU fallback)")
        corr = self._corr
        if isinstance(corr, torch.Tensor):
            d = corr.detach()
            if d.dtype != torch.float32 or not d.is_cuda:
                raise TypeError("corr tensor must be a float32 CUDA tensor")
            d = d.contiguous().view(-1)
            dev = d.device
        else:
            a = np.asarray(corr)
            if a.dtype != np.float32:
                raise TypeError("corr must be float32 (the dtype the GPU search returns), got %s" % a.dtype)
            from .correlator import _resolve_device
            dev = _resolve_device(self._device)               # LOCAL_RANK, then the current device
            d = torch.from_numpy(np.ascontiguousarray(a).reshape(-1)).to(dev)
        n = d.numel()
        with torch.cuda.device(dev):
            s = torch.cuda.current_stream(dev).cuda_stream
            mx = torch.empty(1, dtype=torch.float32, device=dev)
            scratch = torch.empty(1, dtype=torch.int32, device=dev)
            _lib.check(lib.pfb_lcc_max(d.data_ptr(), n, mx.data_ptr(), scratch.data_ptr(), s))
            max_cc = mx.cpu().numpy()[0]                      # np.float32, like corr.max()
            cutoffs = watershed_cutoffs(max_cc, self._steps)
            lowest = np.float32(cutoffs[-1])
            cap = 1 << 16
            while True:
                idx = torch.empty(cap, dtype=torch.int32, device=dev)
                val = torch.empty(cap, dtype=torch.float32, device=dev)
                cnt = torch.empty(1, dtype=torch.int32, device=dev)
                _lib.check(lib.pfb_peak_candidates(d.data_ptr(), n, float(lowest), cap, idx.data_ptr(),
                                                   val.data_ptr(), cnt.data_ptr(), s))
                c = int(cnt.cpu()[0])
                if c <= cap:
                    break
                cap = 1 << int(np.ceil(np.log2(c)))
            return max_cc, cutoffs, idx[:c].cpu().numpy(), val[:c].cpu().numpy()

    def _watershed(self):
        """Positions of high correlation values (analyzer.py:80-95)."""
        shape = tuple(self._corr.shape)
        max_cc, cutoffs, idx, val = self._device_candidates()
        self.last_candidates = int(idx.size)
        if max_cc != max_cc:            # NaN maximum: every comparison fails, no features
            self._positions = set()
            return
        lin = sparse_feature_maxima(idx, val, shape, cutoffs)
        self._positions = set(tuple(int(c) for c in np.unravel_index(i, shape)) for i in lin)

    # ------------------------------------------------------------------ host part (rows of solutions.out)
    def _generate_solutions(self):
        corr = self._corr
        if not isinstance(corr, np.ndarray):
            corr = corr.detach().cpu().numpy()
        rows = []
        for pos in self._positions:
            lcc = corr[pos]
            fishers_z = 0.5 * (np.log(1 + lcc) - np.log(1 - lcc))      # analyzer.py:67
            rel_z = fishers_z / self._z_sigma
            z, y, x = [c * self._voxelspacing + shift for c, shift in zip(pos, self._origin[::-1])]
            rotmat = self._rotmat[int(self._rotmat_ind[pos])]
            rows.append([lcc, fishers_z, rel_z, x, y, z] + list(np.asarray(rotmat).ravel()))
        self._solutions = sorted(rows, key=lambda r: r[0], reverse=True)

    def tofile(self, out='solutions.out'):
        """Same columns and formats as analyzer.py:97-113."""
        if self._solutions is None:
            self._generate_solutions()
        names = '#rank cc Fish-z rel-z x y z a11 a12 a13 a21 a22 a23 a31 a32 a33'.split()
        head = ' '.join(['{:<6s}'] + ['{:>6s}'] * 3 + ['{:>8s}'] * 3 + ['{:>6s}'] * 9) + '\n'
        row = ' '.join(['{:<6d}'] + ['{:6.3f}'] * 3 + ['{:8.3f}'] * 3 + ['{:6.3f}'] * 9) + '\n'
        with open(out, 'w') as f:
            f.write(head.format(*names))
            for n, sol in enumerate(self._solutions):
                f.write(row.format(n + 1, *sol))
