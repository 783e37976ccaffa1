"""``PowerFitter`` -- drop-in for the reference's search wrapper on the ``--gpu`` path.

Mirrors /root/reference/src/powerfit_em/powerfitter.py:50-92: the caller sets
``_rotations``, ``_template``, ``_mask`` (Volume-like objects with ``.array``, or
arrays), optionally ``_queues`` / ``_nproc`` / ``directory``, calls ``scan()`` and
reads ``_lcc`` / ``_rot``.  Here ``scan()`` always runs the CUDA correlator (the
reference's ``_gpu_scan``, :83-92); ``_queues`` may carry CUDA device ordinals, and
``_nproc`` is accepted but ignored -- there is no CPU path in this package.  With
``shard=True`` (opt-in; every rank of the process group must then run the same search) the
rotation list is sharded over the ranks of an initialised ``torch.distributed`` group
exactly like ``_cpu_scan`` shards it over processes (:95-108) and merged on the device
instead of through ``.npy`` files (:146-163).
"""
from os.path import abspath, isdir

import numpy as np

from .correlator import CUDACorrelator


def _array_of(obj):
    return obj.array if hasattr(obj, "array") else np.asarray(obj)


class PowerFitter(object):

    def __init__(self, target, laplace=False, shard=False, group=None, result_rank=None):
        self._target = target
        self._shard = shard
        self._group = group
        self._result_rank = result_rank    # with shard=True: only this rank gets _lcc / _rot (see CUDACorrelator)
        self._rotations = None
        self._template = None
        self._mask = None
        self._queues = None
        self._nproc = 1
        self._directory = abspath('./')
        self._laplace = laplace
        self._batch = 0
        self.pad_to_fused = False      # opt-in: see CUDACorrelator(pad=True)

    @property
    def directory(self):
        return self._directory

    @directory.setter
    def directory(self, directory):
        if isdir(directory):
            self._directory = abspath(directory)
        else:
            raise ValueError("Directory does not exist.")

    def scan(self):
        self._gpu_scan()

    def _gpu_scan(self):
        device = None
        if self._queues:
            q = self._queues[0]
            device = q if isinstance(q, (int, str)) or hasattr(q, "type") else None
        self._corr = CUDACorrelator(_array_of(self._target), device=device, laplace=self._laplace,
                                    batch=self._batch, pad=self.pad_to_fused, shard=self._shard,
                                    group=self._group, result_rank=self._result_rank)
        self._corr.template = _array_of(self._template)
        self._corr.mask = _array_of(self._mask)
        self._corr.rotations = self._rotations
        self._corr.scan()
        self._lcc = self._corr.lcc
        self._rot = self._corr.rot
