"""CPU tests of the host-side logic: library loads and exports the C ABI, packing,
rotation sharding, and the world_size-2 merge over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, golden_inputs


def test_library_builds_and_exports_every_declared_symbol():
    from powerfit_b200 import _lib
    path = _lib.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "powerfit_b200.h")).read()
    declared = set(re.findall(r"\b(pfb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.pfb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pfb_version()


def test_no_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from powerfit_b200 import CUDACorrelator, PowerfitB200Error
    with pytest.raises(PowerfitB200Error):
        CUDACorrelator(np.random.rand(8, 8, 8))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "powerfit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f


def test_packing_order_and_roundtrip():
    from powerfit_b200 import packing as P
    lcc = np.array([0.0, -0.0, 1e-30, -1e-30, 0.5, np.inf, -np.inf, 0.3, 0.3], dtype=np.float32)
    rot = np.array([0, 3, 4, 5, 6, 7, 8, 5, 9])
    key = P.pack(lcc, rot)
    l2, r2 = P.unpack(key)
    assert np.array_equal(l2.view(np.int32), lcc.view(np.int32)) and np.array_equal(r2, rot)
    order = np.argsort(key)
    assert list(lcc[order][:4]) == [-np.inf, np.float32(-1e-30), -0.0, 0.0]
    assert P.pack(np.float32(0.3), 5) > P.pack(np.float32(0.3), 9)       # lower index wins ties
    assert P.pack(np.float32(0.0), 0) == P.BEST_INIT
    assert P.pack(np.float32(0.0), 7) < P.BEST_INIT                      # 0 never replaces the init
    assert P.pack(np.float32(np.nan), 3) == P.BEST_INIT                  # NaN never wins
    assert P.pack(np.float32(1e-38), 2 ** 31 - 2) > P.BEST_INIT


def test_shard_bounds_match_reference_partition(oracle):
    from powerfit_b200 import shard_bounds
    for nrot, world in [(648, 8), (7416, 8), (10, 3), (5, 1), (7, 8)]:
        want = oracle.partition_rotations(nrot, world)
        got = [shard_bounds(nrot, world, r) for r in range(world)]
        assert got == want


def test_packed_merge_equals_reference_combine(oracle):
    from powerfit_b200 import packing as P
    g = load_golden("scan_24_laplace_cw")
    target, template, mask = golden_inputs(g, "scan_24_laplace_cw")
    R = g["rotations"]
    blocks = oracle.partition_rotations(len(R), 3)
    parts, keys = [], []
    for a, b in blocks:
        c = oracle.OracleCorrelator(target, laplace=True)
        c.template, c.mask, c.rotations = template, mask, R[a:b]
        c.scan()
        parts.append((np.nan_to_num(c.lcc), c.rot))
        keys.append(P.pack(c.lcc.astype(np.float32), c.rot + a))
    lcc, rot = oracle.combine_partials(parts, len(R) // 3, target.shape)
    ml, mr = P.unpack(np.maximum.reduce(keys))
    assert np.allclose(ml, lcc, atol=1e-6)
    assert np.array_equal(mr, rot.astype(np.int32))
    assert np.allclose(ml, g["lcc"], atol=1e-6)


WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from powerfit_b200 import packing as P, shard_bounds
from oracle import oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
g = np.load(os.path.join(%(root)r, "tests", "golden", "scan_16x18x20_plain.npz"))
target, template, mask = (g[k].astype(np.float64) for k in ("target", "template", "mask"))
R = g["rotations"]
lo, hi = shard_bounds(len(R), 2, dist.get_rank())
c = O.OracleCorrelator(target); c.template = template; c.mask = mask; c.rotations = R[lo:hi]; c.scan()
key = torch.from_numpy(P.pack(c.lcc.astype(np.float32), c.rot + lo))
root_key = key.clone()
dist.all_reduce(key, op=dist.ReduceOp.MAX)
lcc, rot = P.unpack(key.numpy())
ok = np.allclose(lcc, g["lcc"], atol=1e-6)
decided = (g["lcc"] - g["lcc2"]) > 1e-6
ok = ok and np.array_equal(rot[decided], g["rot"][decided])
# result_rank: a MAX reduce to one rank gives that rank the same merged grid (CUDACorrelator(result_rank=1))
dist.reduce(root_key, dst=1, op=dist.ReduceOp.MAX)
if dist.get_rank() == 1:
    ok = ok and np.array_equal(root_key.numpy(), key.numpy())
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_world_size_2_gloo_merge(tmp_path):
    """The N>1 host path: shard -> per-rank partial -> packed int64 MAX all-reduce."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    codes = [p.wait(timeout=300) for p in procs]
    assert codes == [0, 0]


@pytest.mark.parametrize("name", ["scan_config1_64", "scan_32_plain", "scan_24_laplace_cw",
                                  "scan_config2_128_subset", "rough"])
def test_sparse_feature_maxima_equals_reference_labelling(name):
    """Host half of the Analyzer (N1): connected components on the short above-cutoff list give
    the positions scipy label + maximum_position give on the full grid (reference golden).  The
    device compaction is replaced by numpy here; the GPU test runs the real kernels."""
    from test_oracle import analyzer_case
    from powerfit_b200.analyzer import sparse_feature_maxima, watershed_cutoffs
    lcc, rot, rotations, steps, vs, origin, zs, positions, solutions, _ = analyzer_case(name)
    cut = watershed_cutoffs(lcc.max(), steps)
    assert all(c.dtype == lcc.dtype for c in cut)
    idx = np.nonzero(lcc.ravel() >= cut[-1])[0]
    rng = np.random.default_rng(0)
    perm = rng.permutation(idx.size)                       # the device returns them in any order
    lin = sparse_feature_maxima(idx[perm], lcc.ravel()[idx][perm], lcc.shape, cut)
    got = set(tuple(int(c) for c in np.unravel_index(i, lcc.shape)) for i in lin)
    assert got == positions
    assert idx.size < 0.2 * lcc.size


def test_laplace_operation_order_is_scipys():
    """prep.cu evaluates scipy.ndimage.laplace(mode='wrap') as (d2_z + d2_y) + d2_x with
    d2 = (-2 x[i]) + (x[i-1] + x[i+1]); this pins that order bit for bit against scipy
    (what BaseCorrelator._laplace_filter calls, powerfitter.py:212-215)."""
    from scipy.ndimage import laplace
    rng = np.random.default_rng(5)
    for shape in [(7, 9, 8), (16, 16, 16), (5, 6, 31)]:
        x = rng.normal(size=shape) * 10.0 ** rng.integers(-3, 4, size=shape)
        d2 = lambda a, ax: (-2.0 * a) + (np.roll(a, 1, ax) + np.roll(a, -1, ax))
        assert np.array_equal((d2(x, 0) + d2(x, 1)) + d2(x, 2), laplace(x, mode="wrap"))


def test_wrapped_padding_preserves_offsets():
    """pad_wrapped keeps every voxel at its signed offset from voxel 0 (the template's centre)."""
    from powerfit_b200.correlator import pad_wrapped, pad_target, fused_cube, fused_shape
    rng = np.random.default_rng(3)
    a = rng.random((10, 13, 9))
    p = pad_wrapped(a, 64)
    assert p.sum() == a.sum() and p.shape == (64, 64, 64)
    for idx in [(0, 0, 0), (3, 6, 4), (9, 12, 8), (5, 7, 5), (6, 6, 4)]:
        off = [i if i <= s // 2 else i - s for i, s in zip(idx, a.shape)]
        assert p[tuple(o % 64 for o in off)] == a[idx]
    t = pad_target(a, 64)
    assert np.array_equal(t[:10, :13, :9], a) and t.sum() == a.sum()
    assert fused_cube((40, 52, 46)) == 64 and fused_cube((100, 90, 84)) == 128 and fused_cube((300, 10, 10)) is None
    # per-axis padding: every axis to the next of 32 / 64 / 96 / 128, cubes beyond
    assert fused_shape((28, 40, 27)) == (32, 64, 32) and fused_shape((96, 128, 64)) == (96, 128, 64)
    assert fused_shape((100, 90, 84)) == (128, 96, 96) and fused_shape((130, 20, 20)) == (192, 192, 192)
    assert fused_shape((200, 256, 31)) == (256, 256, 256) and fused_shape((300, 10, 10)) is None
    q = pad_wrapped(a, (32, 64, 32))
    assert q.shape == (32, 64, 32) and np.array_equal(np.sort(q[q != 0]), np.sort(a.ravel()))
    for idx in [(0, 0, 0), (3, 6, 4), (9, 12, 8), (5, 7, 5), (6, 6, 4)]:
        off = [i if i <= s // 2 else i - s for i, s in zip(idx, a.shape)]
        assert q[tuple(o % n for o, n in zip(off, q.shape))] == a[idx]
    assert pad_target(a, (32, 64, 32)).shape == (32, 64, 32)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle port of the reference CPU path on the host cores): one JSON
    line with the keys the driver reads, same metric / unit / workload naming as the CUDA arm."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "config1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "rotations/s (LCC search)" and d["unit"] == "rotations/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "64^3" in d["config"]["workload"]


def test_target_prep_host_steps_match_reference():
    """trim / extend / nearest_multiple2357 of powerfit_b200.target_prep against the reference's
    (tests/golden/make_golden_target_prep.py); the resample step before them needs the GPU."""
    from powerfit_b200 import target_prep as T
    g = load_golden("target_prep")
    assert [T.nearest_multiple2357(n) for n in range(1, 300)] == list(g["smooth"])
    r = g["resampled"]
    vs = float(g["final_vs"])
    sub, origin = T.trim(r, vs, list(g["origin"]), r.max() / 10)
    assert np.array_equal(sub, g["trimmed"]) and np.allclose(origin, g["trim_origin"], rtol=0, atol=0)
    shape = [T.nearest_multiple2357(n) for n in sub.shape]
    assert np.array_equal(T.extend(sub, shape), g["final"])
    e = T.extend(sub, (16, 20, 15))
    assert e.shape == (16, 20, 15) and np.count_nonzero(e) == np.count_nonzero(sub) and np.array_equal(e[:14, :18, :14], sub)
    with pytest.raises(ValueError, match="Cutoff value should be lower than density max."):
        T.trim(r, vs, [0, 0, 0], r.max())


def test_template_work_items_cover_every_rotation_once():
    """(template, rotation block) items of the multi-template search: every (template, rotation) exactly once,
    blocks contiguous like powerfitter.py:95-108, and the same number of items on every rank."""
    from powerfit_b200 import template_work_items
    for T, nrot, world in [(4, 207576, 8), (4, 1000, 2), (4, 7416, 1), (3, 101, 4), (1, 648, 8), (5, 50, 3), (4, 10, 8)]:
        items = template_work_items(T, nrot, world)
        assert len(items) % world == 0
        per_rank = [len(items[r::world]) for r in range(world)]
        assert len(set(per_rank)) == 1
        seen = np.zeros((T, nrot), dtype=np.int64)
        for t, lo, hi in items:
            assert 0 <= lo <= hi <= nrot
            seen[t, lo:hi] += 1
        assert (seen == 1).all()


def test_volume_io_matches_reference_files(tmp_path):
    """N3 file IO: CCP4 / MRC files written by the real reference (tests/golden/make_golden_volume_io.py) parse
    to exactly what the reference's parser returns, and our writer produces the reference's bytes."""
    from powerfit_b200 import volume_io as V
    g = load_golden("volume_io")
    for name in g["names"]:
        name = str(name)
        ext = str(g[name + "_ext"])
        raw = g[name + "_bytes"].tobytes()
        path = tmp_path / (name + "." + ext)
        path.write_bytes(raw)
        dens, vs, origin = V.parse_volume(str(path))
        want = g[name + "_density"]
        assert dens.dtype == want.dtype and np.array_equal(dens, want), name
        assert vs == float(g[name + "_voxelspacing"]) and np.array_equal(np.asarray(origin, dtype=np.float64), g[name + "_origin"])
        with open(path, "rb") as fh:                          # an open file works for every extension here
            d2, vs2, o2 = V.parse_volume(fh)
        assert np.array_equal(d2, want) and vs2 == vs
        vol = V.Volume.fromfile(str(path))
        assert vol.shape == want.shape and np.array_equal(vol.array, want)
        # writer: same bytes as the reference wrote from the same array
        meta = g[name + "_meta"]
        out = tmp_path / ("out_" + name + "." + ext)
        V.Volume(g[name + "_array"], float(meta[0]), tuple(float(v) for v in meta[1:])).tofile(str(out))
        assert out.read_bytes() == raw, name
        # float32 block for the GPU upload
        a, vs3, o3, _ = V.read_map_f32(str(path))
        assert a.dtype == np.float32 and np.array_equal(a, want.astype(np.float32)) and vs3 == vs
    # big-endian file: same values
    name = "f64_mrc"
    raw = bytearray(g[name + "_bytes"].tobytes())
    head = np.frombuffer(bytes(raw[:1024]), dtype=V._header_dtype("<"))[0]
    be = np.zeros(1, dtype=V._header_dtype(">"))[0]
    for k in head.dtype.names:
        be[k] = head[k]
    hb = bytearray(be.tobytes())
    hb[208:216] = b"MAP \x11\x11\x00\x00"
    data = np.frombuffer(bytes(raw[1024:]), dtype="<f4").astype(">f4").tobytes()
    p = tmp_path / "be.mrc"
    p.write_bytes(bytes(hb) + data)
    dens, vs, origin = V.parse_volume(str(p))
    assert np.array_equal(dens, g[name + "_density"]) and vs == float(g[name + "_voxelspacing"])


def test_volume_io_errors(tmp_path):
    """The reference's refusals: unknown extension, bad machine stamp, skewed cell, unequal spacing."""
    from powerfit_b200 import volume_io as V
    g = load_golden("volume_io")
    raw = bytearray(g["f64_mrc_bytes"].tobytes())
    with pytest.raises(ValueError, match="Extension of file is not supported"):
        V.parse_volume(str(tmp_path / "x.txt"))
    bad = bytearray(raw); bad[212] = 0x00
    (tmp_path / "stamp.mrc").write_bytes(bytes(bad))
    with pytest.raises(RuntimeError, match="Endiannes"):
        V.parse_volume(str(tmp_path / "stamp.mrc"))
    import struct
    bad = bytearray(raw); bad[52:56] = struct.pack("<f", 80.0)            # alpha
    (tmp_path / "skew.mrc").write_bytes(bytes(bad))
    with pytest.raises(RuntimeError, match="rectangular"):
        V.parse_volume(str(tmp_path / "skew.mrc"))
    bad = bytearray(raw); bad[40:44] = struct.pack("<f", 99.0)            # xlength
    (tmp_path / "spacing.mrc").write_bytes(bytes(bad))
    with pytest.raises(RuntimeError, match="Voxel spacing"):
        V.parse_volume(str(tmp_path / "spacing.mrc"))
    bad = bytearray(raw); bad[64:68] = struct.pack("<i", 2); bad[68:72] = struct.pack("<i", 1)    # mapc, mapr swapped
    (tmp_path / "order.mrc").write_bytes(bytes(bad))
    with pytest.raises(RuntimeError, match="axis order"):
        V.parse_volume(str(tmp_path / "order.mrc"))
    with pytest.raises(TypeError):
        V.to_mrc(str(tmp_path / "c.mrc"), V.Volume(np.zeros((2, 2, 2), dtype=np.complex64)))
    with pytest.raises(RuntimeError, match="Format is not supported"):
        V.Volume(np.zeros((2, 2, 2))).tofile(str(tmp_path / "c.xplor"))
