"""Target-map preparation before the search, array level.

Mirror of what the reference CLI does between reading the map and building the ``PowerFitter``
(/root/reference/src/powerfit_em/powerfit.py:219-233): ``resample`` to 2 x Nyquist (volume.py:66-72;
the linear zoom runs on the device, ``pyramid.resample``), ``trim`` (volume.py:75-95) and ``extend`` to
the next 2.3.5.7-smooth shape (volume.py:97-118) -- or, on request, to the next cubic grid with a fused
pipeline.  Trim and extend are index arithmetic on the host, exactly as in the reference; maps are
``(array, voxelspacing, origin)`` with the origin in x, y, z order like ``Volume.origin``.
"""
from __future__ import annotations

import numpy as np

from . import pyramid
from .correlator import fused_shape


def is_multiple2357(num):                              # volume.py:111-118
    for multiple in (2, 3, 5, 7):
        while num % multiple == 0:
            num //= multiple
    return num == 1


def nearest_multiple2357(num):                         # volume.py:104-108
    nearest = num
    while not is_multiple2357(nearest):
        nearest += 1
    return nearest


def trim(array, voxelspacing, origin, cutoff, margin=2):
    """volume.py:75-95: cut away the slabs whose maximum does not exceed ``cutoff``, keeping ``margin``."""
    if array.max() <= cutoff:
        raise ValueError("Cutoff value should be lower than density max.")
    extent = []
    for axis in range(array.ndim):
        other = tuple(a for a in range(array.ndim) if a != axis)
        above = np.nonzero(array.max(axis=other) > cutoff)[0]
        low = max(0, int(above[0]) - margin)
        high = min(array.shape[axis], int(above[-1]) + 1 + margin)
        extent.append(slice(low, high))
    sub = array[tuple(extent)]
    new_origin = [o + voxelspacing * ext.start for o, ext in zip(origin, extent[::-1])]
    return sub, new_origin


def extend(array, shape):
    """volume.py:97-102: zero-pad at the high end of every axis."""
    out = np.zeros(tuple(shape), dtype=np.float64)
    out[tuple(slice(s) for s in array.shape)] = array
    return out


def prepare_target(array, voxelspacing, origin, resolution, resampling_rate=2, no_resampling=False,
                   no_trimming=False, trimming_cutoff=None, fused=False, device=None):
    """powerfit.py:219-233.  Returns (array, voxelspacing, origin).  ``fused=True`` extends to the next
    grid with a fused pipeline (every axis up to 32/64/96/128; 192^3 / 256^3 beyond) instead of the next
    2.3.5.7-smooth shape -- a legitimate choice of the CLI's `extend` size, so the search result is what the reference
    computes on the same extended map."""
    array = np.asarray(array, dtype=np.float64)
    origin = list(origin)
    if not no_resampling:
        factor = 2 * resampling_rate * voxelspacing / resolution
        if factor < 0.9:
            array, voxelspacing = pyramid.resample(array, voxelspacing, factor, device=device)
    if not no_trimming:
        if trimming_cutoff is None:
            trimming_cutoff = array.max() / 10
        array, origin = trim(array, voxelspacing, origin, trimming_cutoff)
    shape = [nearest_multiple2357(n) for n in array.shape]
    if fused:
        n = fused_shape(shape)
        if n is not None:
            shape = list(n)
    return extend(array, shape), voxelspacing, origin
