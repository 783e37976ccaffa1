"""powerfit_b200 -- B200-native exhaustive local-cross-correlation search.

A from-scratch replacement for the `--gpu` path of haddocking/powerfit
(`PowerFitter.scan()` / `GPUCorrelator`): hand-written sm_100a CUDA kernels behind a
C ABI (`include/powerfit_b200.h`), driven from Python.  Importing the package does not
need a GPU; constructing a correlator does (there is no CPU fallback).
"""
from ._lib import PowerfitB200Error, build, load  # noqa: F401
from .correlator import CUDACorrelator, MultiTemplateCorrelator, shard_bounds, template_work_items  # noqa: F401
from .powerfitter import PowerFitter  # noqa: F401
from .analyzer import Analyzer  # noqa: F401
from . import shapes, pyramid, target_prep, volume_io  # noqa: F401

__version__ = "0.1.0"
