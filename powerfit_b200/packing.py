"""Host-side twin of the packed best key used on the device (csrc/common.cuh).

key = (orderable(lcc_f32) << 32) | (0xFFFFFFFF - rot), a signed 64-bit integer whose
plain MAX reproduces the reference's merge rule: greater LCC wins, equal LCC keeps the
lower rotation index (powerfitter.py:146-163, 327-330).  Used by the multi-rank merge
(torch.distributed MAX all-reduce on int64) and by the host-logic tests.
"""
import numpy as np

BEST_INIT = np.int64(0x00000000FFFFFFFF)      # (+0.0f, rotation 0)


def orderable(lcc):
    s = np.ascontiguousarray(lcc, dtype=np.float32).view(np.int32)
    return s ^ ((s >> 31) & np.int32(0x7FFFFFFF))


def pack(lcc, rot):
    """Pack float32 LCC and rotation index grids; NaN candidates become the initial key."""
    lcc = np.asarray(lcc, dtype=np.float32)
    hi = orderable(lcc).astype(np.int64) << 32
    lo = (np.int64(0xFFFFFFFF) - np.asarray(rot).astype(np.int64)) & np.int64(0xFFFFFFFF)
    key = hi | lo
    return np.where(np.isnan(lcc), BEST_INIT, key)


def unpack(key):
    key = np.asarray(key, dtype=np.int64)
    k = (key >> 32).astype(np.int32)
    lcc = (k ^ ((k >> 31) & np.int32(0x7FFFFFFF))).view(np.float32)
    rot = (np.int64(0xFFFFFFFF) - (key & np.int64(0xFFFFFFFF))).astype(np.int32)
    return lcc, rot
