"""Density-map file IO for the search: CCP4 / MRC maps in and out.

Same observable behaviour as /root/reference/src/powerfit_em/volume.py:227-497 for what the CLI uses
(`Volume.fromfile(...)` of the target map, `Volume(...).tofile("lcc.mrc")` of the result,
powerfit.py:208-212, 306-308):

* ``parse_volume(fid, fmt)`` -> ``(density, voxelspacing, origin)``: the format comes from the extension
  (``ccp4`` / ``map``: origin = start indices x voxel spacing, volume.py:362-367; ``mrc``: origin = the
  header's origin words, :405-410), the byte order from the machine stamp at byte 212 (:321-331), mode 0 / 1 /
  2 data (:369-378, int16 widened to int32 and float32 to float64, :393-397), non-orthogonal cells and unequal
  voxel spacings rejected with the reference's messages (:292-318).  Like the reference only the standard
  axis order (columns = x, rows = y, sections = z) is read; any other order raises instead of running into the
  reference's unassigned ``self.density`` (:380-391).
* ``to_mrc(fid, volume)`` writes byte for byte what the reference's writer writes (:413-497).
* ``Volume`` carries ``array``, ``voxelspacing``, ``origin`` and the derived ``shape`` / ``dimensions`` /
  ``start`` (:14-55).

What is different is where the voxels land: the data block is read straight into page-locked host memory
(``pinned=True`` and a CUDA build of torch), so the upload to the device that follows is one DMA transfer
without a staging copy; ``read_map_f32`` returns that float32 block itself for callers that feed the GPU.
The header is decoded with one numpy structured dtype instead of a struct format string.
"""
from __future__ import annotations

import os

import numpy as np

_HEADER_BYTES = 1024


def _header_dtype(endian):
    e = endian
    return np.dtype([
        ("nc", e + "i4"), ("nr", e + "i4"), ("ns", e + "i4"), ("mode", e + "i4"),
        ("ncstart", e + "i4"), ("nrstart", e + "i4"), ("nsstart", e + "i4"),
        ("nx", e + "i4"), ("ny", e + "i4"), ("nz", e + "i4"),
        ("xlength", e + "f4"), ("ylength", e + "f4"), ("zlength", e + "f4"),
        ("alpha", e + "f4"), ("beta", e + "f4"), ("gamma", e + "f4"),
        ("mapc", e + "i4"), ("mapr", e + "i4"), ("maps", e + "i4"),
        ("amin", e + "f4"), ("amax", e + "f4"), ("amean", e + "f4"),
        ("ispg", e + "i4"), ("nsymbt", e + "i4"), ("lskflg", e + "i4"),
        ("skwmat", e + "f4", (9,)), ("skwtrn", e + "f4", (3,)), ("extra", e + "f4", (12,)),
        ("origin", e + "f4", (3,)), ("map", "S4"), ("machst", "S4"), ("rms", e + "f4"),
        ("nlabel", e + "i4"), ("label", "S800")])


assert _header_dtype("<").itemsize == _HEADER_BYTES

_MODE_DTYPE = {0: "i1", 1: "i2", 2: "f4"}


def _pinned_empty(count, dtype, pinned):
    """A 1-D numpy array of `count` items; backed by page-locked memory when asked for and possible."""
    if pinned:
        try:
            import torch
            if torch.cuda.is_available():
                buf = torch.empty(count * np.dtype(dtype).itemsize, dtype=torch.uint8).pin_memory()
                return buf.numpy().view(dtype), buf
        except Exception:
            pass
    return np.empty(count, dtype=dtype), None


class MapFile(object):
    """Header and data block of one CCP4 / MRC file."""

    def __init__(self, fid, mrc_origin=False, pinned=True):
        own = isinstance(fid, (str, os.PathLike))
        handle = open(fid, "rb") if own else fid
        if not hasattr(handle, "readinto"):
            raise ValueError("Input should either be a file or filename.")
        try:
            handle.seek(212)
            stamp = handle.read(1)
            if stamp == b"\x44":
                self.endian = "<"
            elif stamp == b"\x11":
                self.endian = ">"
            else:
                raise RuntimeError("Endiannes is not properly set in file. Check the file format.")
            handle.seek(0)
            raw = handle.read(_HEADER_BYTES)
            if len(raw) != _HEADER_BYTES:
                raise RuntimeError("File is shorter than a CCP4/MRC header.")
            h = np.frombuffer(raw, dtype=_header_dtype(self.endian))[0]
            self.header = {k: (h[k].tolist() if h[k].shape else h[k].item()) for k in h.dtype.names}
            self.header["map"] = raw[208:212].decode("latin-1")
            self.header["machst"] = raw[212:216].decode("latin-1")
            self.header["label"] = raw[224:1024].decode("latin-1")
            for name in ("alpha", "beta", "gamma"):
                if abs(self.header[name] - 90) > 1e-3:
                    raise RuntimeError("Only densities in rectangular boxes are supported.")
            self.order = (self.header["mapc"], self.header["mapr"], self.header["maps"])
            spacings = [self.header[a + "length"] / float(self.header["n" + a]) for a in "xyz"]
            mean = sum(spacings) / 3.0
            if any(abs(s - mean) > 1e-4 for s in spacings):
                raise RuntimeError("Voxel spacing is not equal in all directions.")
            self.voxelspacing = spacings[0]
            if mrc_origin:
                self.origin = list(self.header["origin"])
            else:
                start = [self.header[k] for k in ("nsstart", "nrstart", "ncstart")]
                self.origin = np.asarray([start[ax - 1] * self.voxelspacing for ax in self.order])
            if self.order != (1, 2, 3):
                raise RuntimeError("Only the standard axis order (mapc, mapr, maps) = (1, 2, 3) is supported.")
            mode = self.header["mode"]
            if mode not in _MODE_DTYPE:
                raise RuntimeError("Data mode {:} is not supported.".format(mode))
            self.mode = mode
            self.shape = (self.header["nz"], self.header["ny"], self.header["nx"])
            count = int(np.prod(self.shape))
            dtype = np.dtype(self.endian + _MODE_DTYPE[mode])
            # like the reference's np.fromfile right after the header: the symmetry records (nsymbt bytes) are
            # not skipped
            data, self._pin = _pinned_empty(count, dtype, pinned)
            got = handle.readinto(memoryview(data).cast("B"))
            if got != count * dtype.itemsize:
                raise ValueError("cannot reshape array of size {:} into shape {:}".format(
                    got // dtype.itemsize, tuple(self.shape)))
            self.raw = data.reshape(self.shape)
        finally:
            if own:
                handle.close()

    @property
    def density(self):
        """The array the reference returns: float64 for mode 2, int32 for mode 1, int8 for mode 0."""
        if self.mode == 2:
            return self.raw.astype(np.float64)
        if self.mode == 1:
            return self.raw.astype(np.int32)
        return self.raw.astype(np.int8)


def _format_of(fid, fmt):
    name = getattr(fid, "name", fid)
    if fmt is None:
        fmt = os.path.splitext(str(name))[-1][1:]
    return fmt


def parse_volume(fid, fmt=None, pinned=False):
    """volume.py:227-244: ``(density, voxelspacing, origin)`` of a ``.ccp4`` / ``.map`` / ``.mrc`` file."""
    fmt = _format_of(fid, fmt)
    if fmt in ("ccp4", "map"):
        m = MapFile(fid, mrc_origin=False, pinned=pinned)
    elif fmt == "mrc":
        m = MapFile(fid, mrc_origin=True, pinned=pinned)
    elif fmt in ("xplor", "cns"):
        raise ValueError("XPLOR/CNS maps are not supported (the reference's reader for them needs Python 2).")
    else:
        raise ValueError("Extension of file is not supported.")
    return m.density, m.voxelspacing, m.origin


def read_map_f32(fid, fmt=None):
    """The data block as native float32 in page-locked memory (when torch has CUDA), plus voxel spacing and
    origin: what the GPU search uploads.  Mode 2 little-endian files are returned without any copy."""
    fmt = _format_of(fid, fmt)
    if fmt not in ("ccp4", "map", "mrc"):
        raise ValueError("Extension of file is not supported.")
    m = MapFile(fid, mrc_origin=(fmt == "mrc"), pinned=True)
    a = m.raw
    if a.dtype != np.dtype("<f4"):
        out, pin = _pinned_empty(a.size, np.dtype("<f4"), True)
        out = out.reshape(a.shape)
        out[...] = a
        a = out
        m._pin = pin
    a = a.view(np.float32)
    return a, m.voxelspacing, m.origin, m


class Volume(object):
    """volume.py:12-55."""

    @classmethod
    def fromfile(cls, fid, fmt=None):
        array, voxelspacing, origin = parse_volume(fid, fmt)
        return cls(array, voxelspacing, origin)

    def __init__(self, array, voxelspacing=1.0, origin=(0, 0, 0)):
        self.array = array
        self.voxelspacing = voxelspacing
        self.origin = origin

    @property
    def shape(self):
        return self.array.shape

    @property
    def dimensions(self):
        return np.asarray([n * self.voxelspacing for n in self.array.shape][::-1])

    @property
    def start(self):
        return np.asarray([o / self.voxelspacing for o in self.origin])

    def duplicate(self):
        return Volume(self.array.copy(), voxelspacing=self.voxelspacing, origin=self.origin)

    def tofile(self, fid, fmt=None):
        if fmt is None:
            fmt = os.path.splitext(fid)[-1][1:]
        if fmt in ("ccp4", "map", "mrc"):
            to_mrc(fid, self, fmt=fmt)
        else:
            raise RuntimeError("Format is not supported.")


def to_mrc(fid, volume, fmt=None):
    """volume.py:413-497: header (native byte order, machine stamp 0x44 0x41) + data as int8 / int16 / float32."""
    if fmt is None:
        fmt = os.path.splitext(fid)[-1][1:]
    if fmt not in ("ccp4", "mrc", "map"):
        raise ValueError("Format is not recognized. Use ccp4, mrc, or map.")
    kind = volume.array.dtype.name
    if kind == "int8":
        mode = 0
    elif kind in ("int16", "int32"):
        mode = 1
    elif kind in ("float32", "float64"):
        mode = 2
    else:
        raise TypeError("Data type ({:})is not supported.".format(kind))
    nz, ny, nx = volume.shape
    h = np.zeros(1, dtype=_header_dtype("="))[0]
    h["nc"], h["nr"], h["ns"], h["mode"] = nx, ny, nz, mode
    if fmt in ("ccp4", "map"):
        h["ncstart"], h["nrstart"], h["nsstart"] = [int(round(x)) for x in volume.start]
    h["nx"], h["ny"], h["nz"] = nx, ny, nz
    h["xlength"], h["ylength"], h["zlength"] = volume.dimensions
    h["alpha"] = h["beta"] = h["gamma"] = 90.0
    h["mapc"], h["mapr"], h["maps"] = 1, 2, 3
    h["amin"], h["amax"], h["amean"] = volume.array.min(), volume.array.max(), volume.array.mean()
    h["ispg"] = 1
    if fmt == "mrc":
        h["origin"] = np.asarray(volume.origin, dtype=np.float32)
    h["map"] = b"MAP "
    h["machst"] = b"\x44\x41\x00\x00"
    h["rms"] = volume.array.std()
    h["label"] = b" " * 800
    raw = bytearray(h.tobytes())
    raw[212:216] = b"\x44\x41\x00\x00"          # 'S4' drops trailing NULs on assignment: restore them
    with open(fid, "wb") as out:
        out.write(bytes(raw))
        volume.array.astype((np.int8, np.int16, np.float32)[mode]).tofile(out)
