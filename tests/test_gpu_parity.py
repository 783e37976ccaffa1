"""GPU parity tests (run on the B200 box: `pytest -m gpu`).

Everything goes through the C ABI (ctypes) via `CUDACorrelator`.  Tolerances are the
ones BASELINE.json:north_star states: LCC within 1e-4 absolute (FP32 vs the FP64
reference), best-rotation index identical wherever the two best LCCs of a voxel differ
by more than that tolerance.
"""
import numpy as np
import pytest

from conftest import load_golden, golden_inputs

pytestmark = pytest.mark.gpu

LCC_TOL = 1e-4


@pytest.fixture(scope="module")
def pfb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import powerfit_b200
    powerfit_b200.load()            # raises if the CUDA library is missing: no fallback
    return powerfit_b200


def run_scan(pfb, target, template, mask, rotations, laplace, batch=0):
    c = pfb.CUDACorrelator(target, laplace=laplace, batch=batch)
    c.template = template
    c.mask = mask
    c.rotations = rotations
    c.scan()
    assert c.kernel_launches > 0
    return c


def check_against_golden(c, g):
    assert c._rmax == int(g["rmax"]) and float(c._norm_factor) == float(g["norm_factor"])
    lcc, rot = c.lcc, c.rot
    assert lcc.dtype == np.float32 and rot.dtype == np.int32
    assert np.isfinite(lcc).all()
    err = np.abs(lcc - g["lcc"]).max()
    assert err <= LCC_TOL, err
    decided = (g["lcc"] - g["lcc2"]) > LCC_TOL
    assert decided.sum() > 0
    assert np.array_equal(rot[decided], g["rot"][decided])
    lm = np.unpackbits(g["lcc_mask"])[:lcc.size].reshape(lcc.shape).astype(bool)
    assert (lcc[~lm] == 0).all() and (rot[~lm] == 0).all()
    return err


def test_rotate_golden_vectors(pfb):
    """_extensions.rotate_grid3d parity (reference tests/test_extensions.py:10-37 included)."""
    g = load_golden("rotate_vectors")
    done = 0
    for i in range(int(g["n"])):
        grid, R = g["grid_%d" % i], g["rotmat_%d" % i]
        radius, nearest = (int(v) for v in g["meta_%d" % i])
        if radius != min(grid.shape) // 2:
            continue                        # the ABI derives rmax from the shape
        c = pfb.CUDACorrelator(np.abs(grid) + 1.0)
        out = c.rotate(grid, R, nearest=bool(nearest))[0]
        ref = g["out_%d" % i]
        if nearest:
            assert np.array_equal(out, ref.astype(np.float32)), i
        else:
            assert np.allclose(out, ref, rtol=0, atol=2e-6), (i, np.abs(out - ref).max())
        done += 1
    assert done >= 40


def test_rotate_batch_matches_oracle(pfb, oracle):
    from powerfit_b200 import synth
    rng = np.random.default_rng(0)
    for shape in [(32, 32, 32), (20, 24, 28), (15, 21, 25)]:
        grid = rng.normal(size=shape).astype(np.float32).astype(np.float64)
        rots = synth.random_rotations(9, seed=3)
        c = pfb.CUDACorrelator(np.abs(grid) + 1.0)
        for nearest in (False, True):
            out = c.rotate(grid, rots, nearest=nearest)
            for k, R in enumerate(rots):
                ref = np.zeros(shape)
                oracle.rotate_grid3d(grid, R, min(shape) // 2, ref, nearest)
                if nearest:
                    assert np.array_equal(out[k], ref.astype(np.float32))
                else:
                    assert np.allclose(out[k], ref, rtol=0, atol=5e-6)


@pytest.mark.parametrize("shape", [(16, 18, 20), (12, 15, 14), (64, 64, 64), (30, 28, 36), (2, 4, 6),
                                   (49, 25, 27), (128, 128, 128)])
def test_fft3_matches_numpy(pfb, shape):
    rng = np.random.default_rng(1)
    v = (rng.normal(size=(2,) + shape) + 1j * rng.normal(size=(2,) + shape)).astype(np.complex64)
    c = pfb.CUDACorrelator(np.ones(shape))
    out = c.fft3(v)
    ref = np.fft.ifftn(v.astype(np.complex128), axes=(1, 2, 3)) * np.prod(shape)
    scale = np.abs(ref).max()
    assert np.abs(out - ref).max() / scale < 2e-6


@pytest.mark.parametrize("lanes,e", [(4, 4), (4, 8), (4, 12), (4, 24), (8, 8), (8, 16), (8, 24), (16, 16)])
def test_fused_pencil_building_blocks_match_numpy(pfb, lanes, e):
    """Operator-level pin of the register / shared-memory FFT blocks the fused kernels are made of
    (csrc/fft_core.cuh through pfb_pencil_fft): packed pencils, both row transforms with the split radix-2 step,
    and the scalar pencil, for every (lanes, points per lane) geometry the search uses, against numpy."""
    import ctypes
    import torch
    from powerfit_b200 import _lib
    lib = _lib.load()
    n, count = lanes * e, 37                       # 37: a last, partly filled warp
    rng = np.random.default_rng(lanes * 100 + e)
    cplx = lambda *shape: rng.normal(size=shape) + 1j * rng.normal(size=shape)

    def run(kind, packed):
        d_in = torch.from_numpy(np.ascontiguousarray(packed, dtype=np.float32)).cuda()
        d_out = torch.empty_like(d_in)
        _lib.check(lib.pfb_pencil_fft(kind, lanes, e, d_in.data_ptr(), d_out.data_ptr(), count,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return d_out.cpu().numpy().astype(np.float64)

    pack = lambda a, b: np.stack([a.real, b.real, a.imag, b.imag], axis=-1)        # C2 = (re0, re1, im0, im1)
    unpack = lambda o: (o[..., 0] + 1j * o[..., 2], o[..., 1] + 1j * o[..., 3])
    dft = lambda x: np.fft.ifft(x, axis=-1) * x.shape[-1]                          # exp(+2 pi i n k / N)
    tol = lambda ref: 3e-6 * np.abs(ref).max()
    # packed pencil: two independent sequences
    a, b = cplx(count, n), cplx(count, n)
    oa, ob = unpack(run(0, pack(a, b)))
    assert np.abs(oa - dft(a)).max() < tol(dft(a)) and np.abs(ob - dft(b)).max() < tol(dft(b))
    # row transforms of one 2n-point sequence
    x = cplx(count, 2 * n)
    X = dft(x)
    lo, hi = unpack(run(1, pack(x[:, 0::2], x[:, 1::2])))                          # adjacent in -> (X[k], X[k+n])
    assert np.abs(lo - X[:, :n]).max() < tol(X) and np.abs(hi - X[:, n:]).max() < tol(X)
    ev, od = unpack(run(2, pack(x[:, :n], x[:, n:])))                              # split in -> (X[2k], X[2k+1])
    assert np.abs(ev - X[:, 0::2]).max() < tol(X) and np.abs(od - X[:, 1::2]).max() < tol(X)
    if (lanes, e) in ((8, 8), (8, 16), (4, 8), (4, 24)):                           # kernel A's scalar pencils
        o = run(3, np.stack([a.real, a.imag, 0 * a.real, 0 * a.real], axis=-1))    # scalar: float4 = (re, im, -, -)
        assert np.abs(o[..., 0] + 1j * o[..., 1] - dft(a)).max() < tol(dft(a))


def test_lcc_take_best_edge_cases(pfb, oracle):
    """calc_lcc + take-best: zero/negative variance, ties, negative LCC, mask off."""
    shape = (4, 4, 6)
    n = int(np.prod(shape))
    rng = np.random.default_rng(2)
    target = rng.random(shape) + 0.2
    target.reshape(-1)[:7] = 0.0            # below the 5% threshold: lcc_mask off
    c = pfb.CUDACorrelator(target)
    norm = 17.0
    best = None
    ref_lcc = np.zeros(n)
    ref_rot = np.zeros(n)
    lm = c._lcc_mask.reshape(-1)
    for r in range(6):
        gcc = rng.normal(size=n).astype(np.float32)
        ave = rng.normal(size=n).astype(np.float32)
        ave2 = (rng.random(n).astype(np.float32) + 0.5)
        ave2[10] = ave[10] ** 2 / norm * 0.5        # negative variance -> NaN
        ave2[11] = 0.0; ave[11] = 0.0; gcc[11] = abs(gcc[11]) + 0.1     # zero variance -> +inf
        gcc[12], ave[12], ave2[12] = -1.0, 0.5, 1.0
        if r in (2, 4):
            gcc[12] = 0.75                                             # exact tie between r=2 and r=4
        lcc, rot, best = c.lcc_take_best(gcc.reshape(shape), ave.reshape(shape), ave2.reshape(shape),
                                         norm, r, best)
        scan = np.zeros(n)
        with np.errstate(all="ignore"):
            var = (ave2 * np.float32(norm) - ave * ave).astype(np.float32)
            val = (gcc / np.sqrt(var)).astype(np.float32)
        scan[lm != 0] = val[lm != 0]
        ind = scan > ref_lcc
        ref_lcc[ind] = scan[ind]
        ref_rot[ind] = r
    assert np.array_equal(lcc.reshape(-1), ref_lcc.astype(np.float32))
    assert np.array_equal(rot.reshape(-1), ref_rot.astype(np.int32))
    assert np.isinf(lcc.reshape(-1)[11]) and rot.reshape(-1)[12] == 2 and (lcc.reshape(-1)[:7] == 0).all()


def test_contract_errors(pfb):
    rng = np.random.default_rng(3)
    c = pfb.CUDACorrelator(rng.random((8, 10, 12)))
    assert c._target.max() == 1
    with pytest.raises(ValueError):
        c.template = rng.random((3, 3, 3))
    with pytest.raises(ValueError):
        c.mask = np.ones((8, 10, 12))
    c.template = rng.random((8, 10, 12))
    with pytest.raises(ValueError):
        c.mask = np.zeros((8, 10, 12))
    with pytest.raises(ValueError):
        c.scan()
    c.mask = np.ones((8, 10, 12))
    c.rotations = [0] * 27
    assert c.rotations.shape == (3, 3, 3)
    with pytest.raises(ValueError):
        c.rotations = [0] * 3
    with pytest.raises(pfb.PowerfitB200Error):
        pfb.CUDACorrelator(rng.random((11, 8, 8)))          # 11 is not 2.3.5.7-smooth


def test_perfect_fit_gives_unit_lcc(pfb):
    """reference tests/test_powerfitter.py:67-79 through the whole CUDA chain: a template equal to the map,
    an all-ones mask and the identity rotation score LCC = 1 at voxel 0 and nowhere higher.  (The reference
    test bypasses the rotation; here the map lives inside the rmax sphere so that the rotation keeps all of it.)"""
    shape = (10, 12, 14)
    rng = np.random.default_rng(5)
    r = np.array(np.meshgrid(*[np.minimum(np.arange(n), n - np.arange(n)) for n in shape], indexing="ij"))
    inside = (r ** 2).sum(0) <= (min(shape) // 2) ** 2 - 1
    target = np.where(inside, rng.random(shape) + 0.5, 0.0)
    c = pfb.CUDACorrelator(target)
    c._lcc_mask[:] = 1
    c._d_lcc_mask.fill_(1)
    import ctypes
    from powerfit_b200 import _lib
    _lib.check(c._libh.pfb_set_target(c._plan, c._d_target.data_ptr(), c._d_lcc_mask.data_ptr(), c._stream()))
    c.template = target.copy()
    c.mask = inside.astype(np.float64)
    c.rotations = np.eye(3)
    c.scan()
    assert abs(float(c.lcc[0, 0, 0]) - 1.0) < 1e-5, c.lcc[0, 0, 0]
    assert int(np.argmax(c.lcc)) == 0 and c.rot[0, 0, 0] == 0
    assert c.lcc.max() <= 1 + 1e-5


SMALL = ["scan_16x18x20_plain", "scan_12x15x14_laplace", "scan_24_laplace_cw", "scan_32_plain"]


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("batch", [0, 2, 5])
def test_scan_small_golden(pfb, name, batch):
    g = load_golden(name)
    target, template, mask = golden_inputs(g, name)
    c = run_scan(pfb, target, template, mask, g["rotations"], bool(g["laplace"]), batch=batch)
    assert np.allclose(c._template, g["prepped_template"], atol=1e-5)
    check_against_golden(c, g)


def test_scan_small_vs_live_oracle(pfb, oracle):
    from powerfit_b200 import synth
    case = synth.make_case(shape=(20, 24, 18), voxelspacing=3.0, resolution=9.0, n_res=30, rg=6.0,
                           n_copies=2, seed=9, core_weighted=True)
    rots = synth.random_rotations(13, seed=2)          # odd count: last pair half empty
    o = oracle.OracleCorrelator(case.target, laplace=True)
    o.template, o.mask, o.rotations = case.template, case.mask, rots
    o.scan(track_second=True)
    c = run_scan(pfb, case.target, case.template, case.mask, rots, True)
    ref = np.nan_to_num(o.lcc)
    assert np.abs(c.lcc - ref).max() <= LCC_TOL
    decided = (ref - o._lcc2) > LCC_TOL
    assert np.array_equal(c.rot[decided], o.rot[decided].astype(np.int32))


def test_scan_config1_full_golden(pfb):
    """BASELINE config 1: 64^3, 648 rotations, against the reference CPU path's result."""
    g = load_golden("scan_config1_64")
    target, template, mask = golden_inputs(g, "scan_config1_64")
    c = run_scan(pfb, target, template, mask, g["rotations"], False)
    err = check_against_golden(c, g)
    # the best solution (what solutions.out ranks first) is the same voxel and rotation
    assert np.argmax(c.lcc) == np.argmax(g["lcc"])
    assert c.rot.reshape(-1)[np.argmax(c.lcc)] == g["rot"].reshape(-1)[np.argmax(g["lcc"])]
    print("config1 max |dLCC| = %.3g" % err)


@pytest.mark.parametrize("name", ["scan_config2_128_subset", "scan_config3_128_cw_subset"])
def test_scan_128_subset_golden(pfb, name):
    """BASELINE configs 2/3 (128^3, Laplace / core-weighted) on a rotation subset."""
    g = load_golden(name)
    target, template, mask = golden_inputs(g, name)
    c = run_scan(pfb, target, template, mask, g["rotations"], bool(g["laplace"]))
    check_against_golden(c, g)


def solutions_agree(rows, want, top, tol=LCC_TOL):
    """north_star acceptance rule for solutions.out: the top-N solutions have identical positions and rotations.
    Rows whose scores differ by less than the LCC tolerance may swap ranks, so every reference row is looked up
    by position; its rotation matrix must be identical and its score within the tolerance."""
    rows, want = np.asarray(rows, dtype=np.float64), np.asarray(want, dtype=np.float64)
    got = {tuple(r[3:6]): r for r in rows}
    for k, w in enumerate(want[:top]):
        assert tuple(w[3:6]) in got, (k, w[:6])
        r = got[tuple(w[3:6])]
        assert abs(r[0] - w[0]) <= tol, (k, r[0], w[0])
        assert np.array_equal(r[6:], w[6:]), (k, r[6:], w[6:])
    # same order wherever neighbouring reference scores are further apart than the tolerance
    n = min(top, len(want), len(rows))
    for k in range(n):
        lo_ok = k == 0 or want[k - 1][0] - want[k][0] > 2 * tol
        hi_ok = k + 1 >= len(want) or want[k][0] - want[k + 1][0] > 2 * tol
        if lo_ok and hi_ok:
            assert tuple(rows[k][3:6]) == tuple(want[k][3:6]), (k, rows[k][:6], want[k][:6])


def test_gpu_scan_to_solutions_config1(pfb):
    """GPU scan -> Analyzer against reference scan -> reference Analyzer (BASELINE config 1, all 648 rotations):
    solutions.out has the same positions and rotations (tests/golden/make_golden_analyzer.py)."""
    from test_oracle import analyzer_case
    from powerfit_b200.analyzer import Analyzer
    g = load_golden("scan_config1_64")
    target, template, mask = golden_inputs(g, "scan_config1_64")
    c = run_scan(pfb, target, template, mask, g["rotations"], False)
    lcc, rot, rotations, steps, vs, origin, zs, positions, solutions, text = analyzer_case("scan_config1_64")
    a = Analyzer(c.lcc, rotations, c.rot, steps=steps, voxelspacing=vs, origin=origin, z_sigma=zs)
    solutions_agree(a.solutions, solutions, top=min(100, len(solutions)))


def test_scan_config2_full_golden(pfb, tmp_path):
    """BASELINE configs[1] in full: 128^3, Laplace, the whole 10 degree set (7416 rotations), against the
    reference CPU path run in the build container (tests/golden/make_golden_full128.py): LCC within 1e-4 and
    rotation index identical where decided on the stored z planes, the global maximum, and solutions.out --
    GPU scan -> Analyzer against reference scan -> reference Analyzer."""
    from powerfit_b200.analyzer import Analyzer
    g = load_golden("scan_config2_128_full")
    target, template, mask = golden_inputs(g, "scan_config2_128_full")
    c = run_scan(pfb, target, template, mask, g["rotations"], True)
    assert float(c._norm_factor) == float(g["norm_factor"])
    planes = g["planes"]
    lcc, rot = c.lcc[planes], c.rot[planes]
    err = np.abs(lcc - g["lcc"]).max()
    assert err <= LCC_TOL, err
    decided = (g["lcc"] - g["lcc2"]) > LCC_TOL
    assert decided.sum() > 1000
    assert np.array_equal(rot[decided], g["rot"][decided])
    lm = np.unpackbits(g["lcc_mask"])[:lcc.size].reshape(lcc.shape).astype(bool)
    assert (lcc[~lm] == 0).all() and (rot[~lm] == 0).all()
    assert tuple(np.unravel_index(np.argmax(c.lcc), c.lcc.shape)) == tuple(int(v) for v in g["argmax"])
    assert abs(float(c.lcc.max()) - float(g["lcc64_max"])) <= LCC_TOL
    assert abs(float(c.lcc.astype(np.float64).sum()) - float(g["lcc_sum"])) <= 1e-5 * float(g["lcc_sum"])
    steps, vs, ox, oy, oz, zs = (float(v) for v in g["analyzer_params"])
    a = Analyzer(c.lcc, g["rotations"], c.rot, steps=int(steps), voxelspacing=vs, origin=(ox, oy, oz), z_sigma=zs)
    solutions_agree(a.solutions, g["solutions"], top=min(100, len(g["solutions"])))
    out = tmp_path / "solutions.out"
    a.tofile(str(out))
    got, want = out.read_text().splitlines(), str(g["solutions_text"]).splitlines()
    assert got[0] == want[0]
    print("config2 full: max |dLCC| = %.3g, %d solutions (reference %d), identical text rows: %d of %d"
          % (err, len(a.solutions), len(g["solutions"]), sum(x == y for x, y in zip(got, want)), len(want)))


def test_shards_merge_to_single_pass(pfb):
    """Rotation blocks scanned separately (as ranks would) and merged with the packed MAX
    equal one pass over the whole list; rescanning is idempotent."""
    import ctypes, torch
    from powerfit_b200 import _lib, shard_bounds
    g = load_golden("scan_32_plain")
    target, template, mask = golden_inputs(g, "scan_32_plain")
    R = g["rotations"]
    c = run_scan(pfb, target, template, mask, R, False)
    full = c.scan_device().clone()
    parts = []
    for r in range(3):
        lo, hi = shard_bounds(len(R), 3, r)
        parts.append(c.scan_device(lo, hi).clone())
    merged = parts[0].clone()
    lib = _lib.load()
    for p in parts[1:]:
        _lib.check(lib.pfb_merge_best(c._plan, merged.data_ptr(), p.data_ptr(), c._stream()))
    assert torch.equal(merged, full)
    assert torch.equal(torch.maximum(torch.maximum(parts[0], parts[1]), parts[2]), full)
    again = c.scan_device(reset=False)
    assert torch.equal(again, full)


def test_full_size_properties_128(pfb):
    """Size-independent checks at BASELINE's 128^3: a template cut from the map itself
    scores LCC ~ 1 at its true position with the identity rotation, and scaling the map
    leaves the LCC unchanged (the score is normalised)."""
    from powerfit_b200 import synth
    case = synth.config2(seed=1)
    n = 128
    rng = np.random.default_rng(0)
    # map = template shifted by a known vector (+ nothing else): perfect fit
    shift = (17, 90, 41)
    target = np.roll(case.template, shift, axis=(0, 1, 2)) + case.template.max() * (0.1 + 0.01 * rng.random((n, n, n)))
    rots = synth.random_rotations(4, seed=1)            # rots[0] is the identity
    c = run_scan(pfb, target, case.template, case.mask, rots, False)
    peak = np.unravel_index(np.argmax(c.lcc), c.lcc.shape)
    assert peak == shift and abs(c.lcc[peak] - 1) < 2e-2 and c.rot[peak] == 0
    c2 = run_scan(pfb, 7.5 * target, case.template, case.mask, rots, False)
    assert np.abs(c2.lcc - c.lcc).max() < 1e-4 and np.array_equal(c2.rot, c.rot)


@pytest.mark.parametrize("shape,cw,laplace", [((32, 32, 32), False, True), ((24, 30, 28), True, True),
                                              ((64, 64, 64), True, False), ((128, 128, 128), False, True)])
def test_device_prep_equals_host_prep(pfb, shape, cw, laplace):
    """pfb_prepare_target / pfb_prepare_template (FP64 kernels) against the
    reference's numpy/scipy formulas (powerfitter.py:169-220): target, lcc_mask, N and the binary flag
    bit for bit; the z-scored template to one float32 ulp (the masked sums run in another order)."""
    from powerfit_b200 import synth
    case = synth.make_case(shape=shape, voxelspacing=3.0, resolution=9.0, n_res=60, rg=8.0, n_copies=2,
                           seed=31, core_weighted=cw)
    rots = synth.random_rotations(6, seed=2)
    res = {}
    for prep in ("host", "device"):
        c = pfb.CUDACorrelator(case.target, laplace=laplace, prep=prep)
        c.template, c.mask, c.rotations = case.template, case.mask, rots
        c.scan()
        res[prep] = (c._d_target.cpu().numpy(), c._lcc_mask.copy(), c._d_template.cpu().numpy(),
                     c._d_mask.cpu().numpy(), int(c._norm_factor), c._mask_binary, c.lcc.copy(), c.rot.copy())
    h, d = res["host"], res["device"]
    assert np.array_equal(h[0], d[0]) and np.array_equal(h[1], d[1])
    assert np.array_equal(h[3], d[3]) and h[4] == d[4] and h[5] == d[5] == (not cw)
    same = h[2] == d[2]
    assert same.mean() > 0.9999, same.mean()
    assert np.all(np.abs(h[2] - d[2]) <= 2.4e-7 * np.maximum(np.abs(h[2]), 1e-30))
    assert np.abs(h[6] - d[6]).max() < 1e-6
    assert (h[7] == d[7]).mean() > 0.9999


def test_several_templates_on_one_correlator(pfb):
    """BASELINE configs[4] fits several sub-units into one map: re-setting template and mask on the same
    correlator (the map spectra are kept) gives exactly what a fresh correlator gives."""
    from powerfit_b200 import synth
    case = synth.make_case(n=64, voxelspacing=3.0, resolution=9.0, n_res=120, rg=12.0, n_copies=3, seed=41)
    other = synth.make_case(n=64, voxelspacing=3.0, resolution=9.0, n_res=70, rg=9.0, n_copies=1, seed=42,
                            core_weighted=True)
    rots = synth.random_rotations(9, seed=8)
    shared = pfb.CUDACorrelator(case.target, laplace=True)
    for tmpl, mask in ((case.template, case.mask), (other.template, other.mask), (case.template, case.mask)):
        shared.template, shared.mask, shared.rotations = tmpl, mask, rots
        shared.scan()
        fresh = run_scan(pfb, case.target, tmpl, mask, rots, True)
        assert np.array_equal(shared.lcc, fresh.lcc) and np.array_equal(shared.rot, fresh.rot)


def test_multi_template_slots_equal_fresh_correlators(pfb, oracle):
    """BASELINE configs[4] (a batch of distinct templates against one map): MultiTemplateCorrelator keeps one
    template slot per sub-unit on ONE plan (shared FT(map), FT(map^2), work buffers) and scan_all() gives, for every
    template, exactly what a fresh single-template correlator gives -- on a fused grid and on an any-shape grid,
    where it is also checked against the oracle."""
    from powerfit_b200 import MultiTemplateCorrelator, synth
    for shape, laplace in (((64, 64, 64), True), ((20, 24, 18), False)):
        kw = dict(shape=shape, voxelspacing=3.0, resolution=9.0)
        big = shape[0] >= 64
        cases = [synth.make_case(n_res=120 if big else 30, rg=12.0 if big else 6.0, n_copies=3, seed=41, **kw),
                 synth.make_case(n_res=70 if big else 20, rg=9.0 if big else 5.0, n_copies=1, seed=42, core_weighted=True, **kw),
                 synth.make_case(n_res=40 if big else 12, rg=7.0 if big else 4.0, n_copies=1, seed=43, **kw)]
        target = cases[0].target
        rots = synth.random_rotations(11, seed=8)
        m = MultiTemplateCorrelator(target, len(cases), laplace=laplace)
        for i in (2, 0, 1):                                           # any filling order
            m.set_template(i, cases[i].template, cases[i].mask)
        m.rotations = rots
        m.scan_all()
        assert m.last_scan_rotations == len(cases) * len(rots)
        for i, cs in enumerate(cases):
            fresh = run_scan(pfb, target, cs.template, cs.mask, rots, laplace)
            assert np.array_equal(m.lccs[i], fresh.lcc) and np.array_equal(m.rots[i], fresh.rot), i
            if not big:
                o = oracle.OracleCorrelator(target, laplace=laplace)
                o.template, o.mask, o.rotations = cs.template, cs.mask, rots
                o.scan(track_second=True)
                ref = np.nan_to_num(o.lcc)
                assert np.abs(m.lccs[i] - ref).max() <= LCC_TOL
                decided = (ref - o._lcc2) > LCC_TOL
                assert np.array_equal(m.rots[i][decided], o.rot[decided].astype(np.int32))
        # a slot stays valid after others were used: the plain single-template call on the selected slot
        m.select(1)
        m.scan()
        assert np.array_equal(m.lcc, m.lccs[1]) and np.array_equal(m.rot, m.rots[1])
        with pytest.raises(ValueError):
            m.select(len(cases))


def test_padding_to_a_fused_grid(pfb):
    """pad=True: an arbitrary (CLI-like) shape is searched on the next fused grid (every axis rounded up to 32 / 64 /
    96 / 128); the result equals the search on explicitly padded inputs exactly, and the unpadded any-shape search
    away from the box faces."""
    from powerfit_b200 import synth
    from powerfit_b200.correlator import pad_target, pad_wrapped, fused_shape
    shape = (40, 70, 28)
    big = fused_shape(shape)
    assert big == (64, 96, 32)
    case = synth.make_case(shape=shape, voxelspacing=3.0, resolution=9.0, n_res=40, rg=6.0, n_copies=2, seed=51)
    rots = synth.random_rotations(9, seed=5)
    c = pfb.CUDACorrelator(case.target, laplace=False, pad=True)
    c.template, c.mask, c.rotations = case.template, case.mask, rots
    c.scan()
    assert c.plan_info(6) == 1 and c.lcc.shape == shape and c.rot.shape == shape
    e = run_scan(pfb, pad_target(case.target, big), pad_wrapped(case.template, big), pad_wrapped(case.mask, big), rots, False)
    assert e.plan_info(6) == 1
    assert np.array_equal(c.lcc, e.lcc[:40, :70, :28]) and np.array_equal(c.rot, e.rot[:40, :70, :28])
    g = run_scan(pfb, case.target, case.template, case.mask, rots, False)          # any-shape pipeline, periodic box
    assert g.plan_info(6) == 0
    inner = (slice(10, 30), slice(10, 60), slice(10, 18))
    assert np.abs(c.lcc[inner] - g.lcc[inner]).max() < 1e-4


def test_batch_shrinks_when_memory_is_short(pfb):
    """A batch whose work buffers cannot be allocated is halved until they fit; results are unaffected."""
    g = load_golden("scan_32_plain")
    target, template, mask = golden_inputs(g, "scan_32_plain")
    big = np.zeros((64, 64, 64))
    big[:32, :32, :32] = target                        # 64^3 so that the request below is ~630 GB
    c = pfb.CUDACorrelator(big, batch=100000)
    assert 2 <= c.plan_info(4) < 100000
    c2 = run_scan(pfb, target, template, mask, g["rotations"], False, batch=100000)
    check_against_golden(c2, g)


def test_device_prep_contract_errors(pfb):
    t = np.random.default_rng(0).random((12, 12, 12))
    c = pfb.CUDACorrelator(t)
    c.template = t
    with pytest.raises(ValueError, match="Zero-filled mask is not allowed."):
        c.mask = np.zeros_like(t)


def test_search_host_entry_point(pfb):
    """The all-host-buffers C-ABI call gives the same grids as the correlator."""
    import ctypes
    from powerfit_b200 import _lib
    g = load_golden("scan_16x18x20_plain")
    target, template, mask = golden_inputs(g, "scan_16x18x20_plain")
    c = run_scan(pfb, target, template, mask, g["rotations"], False)
    lib = _lib.load()
    V = target.size
    t32 = np.ascontiguousarray(c._target, dtype=np.float32)
    tm32 = np.ascontiguousarray(c._template, dtype=np.float32)
    m32 = np.ascontiguousarray(c._mask, dtype=np.float32)
    lm = np.ascontiguousarray(c._lcc_mask)
    R = np.ascontiguousarray(g["rotations"], dtype=np.float64)
    lcc = np.empty(V, dtype=np.float32)
    rot = np.empty(V, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(lib.pfb_search_host(c._plan, p(t32), p(lm), p(tm32), p(m32), float(c._norm_factor), 1,
                                   p(R), len(R), 0, p(lcc), p(rot)))
    assert np.array_equal(lcc.reshape(target.shape), c.lcc) and np.array_equal(rot.reshape(target.shape), c.rot)


@pytest.mark.parametrize("n,cw,laplace", [(64, False, False), (64, True, True), (128, True, False)])
def test_fused_path_equals_generic_path(pfb, monkeypatch, n, cw, laplace):
    """The fused 3-kernel pipeline (cubic 64/128), with and without support pruning,
    against the any-shape generic pipeline on the same inputs."""
    from powerfit_b200 import synth
    case = synth.make_case(n=n, voxelspacing=2.0 if n == 64 else 2.8, resolution=8.0, n_res=150, rg=12.0,
                           n_copies=3, seed=21, core_weighted=cw)
    rots = synth.random_rotations(7, seed=4)
    results = {}
    for mode, env in [("generic", {"PFB_FUSED": "0"}), ("fused", {"PFB_FUSED": "1"}),
                      ("fused_noprune", {"PFB_FUSED": "1", "PFB_NO_PRUNE": "1"})]:
        for k in ("PFB_FUSED", "PFB_NO_PRUNE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c = run_scan(pfb, case.target, case.template, case.mask, rots, laplace, batch=4)
        assert c.plan_info(6) == (0 if mode == "generic" else 1)
        results[mode] = (c.lcc.copy(), c.rot.copy())
    g_lcc, g_rot = results["generic"]
    for mode in ("fused", "fused_noprune"):
        lcc, rot = results[mode]
        assert np.abs(lcc - g_lcc).max() < 2e-5, mode
        same = rot == g_rot
        assert same.mean() > 0.999, (mode, same.mean())
        assert np.abs(lcc - g_lcc)[~same].max(initial=0) < 2e-5
    assert np.array_equal(results["fused"][0], results["fused_noprune"][0])
    assert np.array_equal(results["fused"][1], results["fused_noprune"][1])


MIXED = [((64, 128, 64), False, False), ((128, 64, 128), True, True), ((64, 64, 128), True, False),
         ((128, 128, 64), False, True), ((96, 128, 64), False, False), ((32, 96, 96), True, True),
         ((96, 96, 96), False, True), ((32, 32, 32), True, False), ((128, 32, 96), False, False),
         ((64, 96, 32), True, True), ((32, 64, 128), False, True), ((96, 32, 64), True, False)]


@pytest.mark.parametrize("shape,cw,laplace", MIXED)
def test_mixed_axis_fused_path_equals_generic_path(pfb, monkeypatch, shape, cw, laplace):
    """Per-axis fused pipeline (any mix of 32 / 64 / 96 / 128 voxels per axis: 4-lane and 8-lane pencils, radix-3
    register DFTs at 96, with and without the TMEM stash / TMA staging / TMA-fed kernel C) against the any-shape
    generic pipeline on the same inputs, with and without support pruning, odd rotation count."""
    from powerfit_b200 import synth
    small = min(shape)
    case = synth.make_case(shape=shape, voxelspacing=3.0, resolution=9.0, n_res=40 if small == 32 else 120,
                           rg=5.0 if small == 32 else 10.0, n_copies=3, seed=sum(shape), core_weighted=cw)
    rots = synth.random_rotations(7, seed=4)
    results = {}
    for mode, env in [("generic", {"PFB_FUSED": "0"}), ("fused", {"PFB_FUSED": "1"}),
                      ("fused_noprune", {"PFB_FUSED": "1", "PFB_NO_PRUNE": "1"})]:
        for k in ("PFB_FUSED", "PFB_NO_PRUNE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c = run_scan(pfb, case.target, case.template, case.mask, rots, laplace, batch=4)
        assert c.plan_info(6) == (0 if mode == "generic" else 1)
        results[mode] = (c.lcc.copy(), c.rot.copy())
    g_lcc, g_rot = results["generic"]
    assert (g_lcc > 0).sum() > 100
    for mode in ("fused", "fused_noprune"):
        lcc, rot = results[mode]
        assert np.abs(lcc - g_lcc).max() < 2e-5, mode
        same = rot == g_rot
        assert same.mean() > 0.999, (mode, same.mean())
        assert np.abs(lcc - g_lcc)[~same].max(initial=0) < 2e-5
    assert np.array_equal(results["fused"][0], results["fused_noprune"][0])
    assert np.array_equal(results["fused"][1], results["fused_noprune"][1])


@pytest.mark.parametrize("name", ["scan_96x128x64_laplace", "scan_32x64x96_cw"])
def test_scan_mixed_axis_golden(pfb, name):
    """Reference CPU search (real reference, tests/golden/make_golden.py) on CLI-like non-cubic grids that take the
    per-axis fused pipeline."""
    g = load_golden(name)
    target, template, mask = golden_inputs(g, name)
    c = run_scan(pfb, target, template, mask, g["rotations"], bool(g["laplace"]))
    assert c.plan_info(6) == 1
    check_against_golden(c, g)


@pytest.mark.parametrize("n,cw,laplace,count", [(256, True, True, 5), (256, False, False, 4), (192, True, True, 5),
                                                 (192, False, False, 6)])
def test_class_path_equals_generic_path(pfb, monkeypatch, n, cw, laplace, count):
    """The class-decimated kernels (fused_cls.cu: 192^3 and 256^3) against the any-shape generic pipeline,
    with and without support pruning, odd and even rotation counts."""
    from powerfit_b200 import synth
    case = synth.make_case(n=n, voxelspacing=2.8, resolution=9.0, n_res=200, rg=14.0,
                           n_copies=3, seed=23, core_weighted=cw)
    rots = synth.random_rotations(count, seed=6)
    results = {}
    for mode, env in [("generic", {"PFB_FUSED": "0"}), ("cls", {}), ("cls_noprune", {"PFB_NO_PRUNE": "1"})]:
        for k in ("PFB_FUSED", "PFB_NO_PRUNE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c = run_scan(pfb, case.target, case.template, case.mask, rots, laplace, batch=4)
        assert c.plan_info(6) == (0 if mode == "generic" else 1)
        assert c.plan_info(9) == (0 if mode == "generic" else 1)
        results[mode] = (c.lcc.copy(), c.rot.copy())
    g_lcc, g_rot = results["generic"]
    for mode in ("cls", "cls_noprune"):
        lcc, rot = results[mode]
        assert np.abs(lcc - g_lcc).max() < 2e-5, mode
        same = rot == g_rot
        assert same.mean() > 0.999, (mode, same.mean())
        assert np.abs(lcc - g_lcc)[~same].max(initial=0) < 2e-5
    assert np.array_equal(results["cls"][0], results["cls_noprune"][0])
    assert np.array_equal(results["cls"][1], results["cls_noprune"][1])


@pytest.mark.parametrize("name,maker,laplace", [("scan_config4_256_subset", "config4", True),
                                                 ("scan_config5_192_subset", "config5", False)])
def test_scan_class_path_subset_golden(pfb, name, maker, laplace):
    """BASELINE configs 4 and 5 shapes (256^3 Laplace + core-weighted, 192^3 plain) on 6 rotations of the
    4.71 degree set plus two true poses, against eight z planes of the reference CPU path's result."""
    from powerfit_b200 import synth
    g = load_golden(name)
    case = getattr(synth, maker)(seed=int(g["seed"]))
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    c = run_scan(pfb, f32(case.target), f32(case.template), f32(case.mask), g["rotations"], laplace)
    assert c.plan_info(6) == 1 and c.plan_info(9) == 1
    assert c._rmax == int(g["rmax"]) and float(c._norm_factor) == float(g["norm_factor"])
    planes = g["planes"]
    lcc, rot = c.lcc[planes], c.rot[planes]
    assert np.isfinite(lcc).all()
    err = np.abs(lcc - g["lcc"]).max()
    assert err <= LCC_TOL, err
    decided = (g["lcc"] - g["lcc2"]) > LCC_TOL
    assert decided.sum() > 1000
    assert np.array_equal(rot[decided], g["rot"][decided])
    lm = np.unpackbits(g["lcc_mask"])[:lcc.size].reshape(lcc.shape).astype(bool)
    assert (lcc[~lm] == 0).all() and (rot[~lm] == 0).all()
    assert tuple(np.unravel_index(np.argmax(c.lcc), c.lcc.shape)) == tuple(int(v) for v in g["argmax"])
    assert abs(float(c.lcc.max()) - float(g["lcc64_max"])) <= LCC_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["scan_config1_64", "scan_32_plain", "scan_24_laplace_cw",
                                  "scan_config2_128_subset", "rough"])
def test_analyzer_matches_reference_solutions(pfb, name, tmp_path):
    """N1: device max + compaction, host labelling -> the reference Analyzer's positions, rows and
    solutions.out text (golden generated by the real reference, tests/golden/make_golden_analyzer.py)."""
    from test_oracle import analyzer_case
    from powerfit_b200.analyzer import Analyzer
    lcc, rot, rotations, steps, vs, origin, zs, positions, solutions, text = analyzer_case(name)
    a = Analyzer(lcc, rotations, rot, steps=steps, voxelspacing=vs, origin=origin, z_sigma=zs)
    assert a._positions == positions
    rows = np.array(a.solutions, dtype=np.float64)
    assert rows.shape == solutions.shape and np.array_equal(rows[:, 0], solutions[:, 0])
    assert np.allclose(rows, solutions, rtol=0, atol=1e-12)
    out = tmp_path / "solutions.out"
    a.tofile(str(out))
    assert out.read_text() == text
    assert 0 < a.last_candidates < 0.2 * lcc.size


@pytest.mark.gpu
def test_analyzer_on_device_tensor_and_search_result(pfb, oracle):
    """Whole hand-off: search on the GPU, analyse the device-resident LCC grid, compare the
    solutions with the CPU restatement run on the same grids."""
    import torch
    from powerfit_b200 import synth
    from powerfit_b200.analyzer import Analyzer
    case = synth.make_case(n=32, voxelspacing=2.0, resolution=8.0, n_res=80, rg=8.0, n_copies=2, seed=9)
    rots = synth.random_rotations(40, seed=2)
    c = run_scan(pfb, case.target, case.template, case.mask, rots, False, batch=8)
    dev = torch.from_numpy(c.lcc).cuda()
    a = Analyzer(dev, rots, c.rot, voxelspacing=2.0, origin=(1.0, 2.0, 3.0), z_sigma=0.05)
    pos = oracle.watershed_positions(c.lcc, 5)
    assert a._positions == pos
    want = np.array(oracle.solution_rows(c.lcc, rots, c.rot, pos, 2.0, (1.0, 2.0, 3.0), 0.05), dtype=np.float64)
    got = np.array(a.solutions, dtype=np.float64)
    assert np.allclose(got, want, rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", ["small", "tiny_box", "config1"])
def test_shapes_match_reference_golden(pfb, oracle, name):
    """N2 on the device (pfb_blur_points, pfb_dilate_points, pfb_core_indices behind
    powerfit_b200.shapes) against grids produced by the real reference: masks and core indices
    exactly, densities to 1e-13 of their maximum (exp() implementations differ in the last bit)."""
    from test_oracle import shapes_case
    from powerfit_b200 import shapes
    shape, vs, origin, res, xyz, weights, radii, vol, mask, mask_r, core = shapes_case(name)
    grid = (shape, vs, origin)
    t = shapes.structure_to_shape_like(grid, xyz, resolution=res, weights=weights, shape="vol")
    assert t.dtype == np.float64 and t.shape == shape
    assert np.abs(t - vol).max() <= 1e-13 * vol.max()
    assert np.array_equal(t == 0, vol == 0)
    m = shapes.structure_to_shape_like(grid, xyz, resolution=res, shape="mask")
    assert np.array_equal(m, mask)
    m2 = shapes.structure_to_shape_like(grid, xyz, resolution=res, radii=radii.copy(), shape="mask")
    assert np.array_equal(m2, mask_r)
    assert np.array_equal(shapes.determine_core_indices(m), core)
    with pytest.raises(ValueError, match="weights array is of incorrect size"):
        shapes.structure_to_shape_like(grid, xyz, resolution=res, weights=weights[:-1], shape="vol")


def test_shapes_feed_the_search(pfb, oracle):
    """Template and core-weighted mask synthesised on the device give the same search result as the
    ones synthesised by the oracle's restatement of the reference code."""
    from powerfit_b200 import shapes, synth
    n, vs, res = 32, 3.0, 9.0
    xyz = synth.random_walk_trace(80, 9.0, 6).T.copy()
    w = np.full(80, 6.0)
    case = synth.make_case(n=n, voxelspacing=vs, resolution=res, n_res=80, rg=9.0, n_copies=2, seed=6)
    grid = ((n, n, n), vs, (0.0, 0.0, 0.0))
    t_d = shapes.structure_to_shape_like(grid, xyz, resolution=res, weights=w, shape="vol")
    m_d = shapes.determine_core_indices(shapes.structure_to_shape_like(grid, xyz, resolution=res, shape="mask"))
    t_o = oracle.structure_to_shape_like((n, n, n), vs, (0, 0, 0), xyz, res, weights=w, kind="vol")
    m_o = oracle.determine_core_indices(oracle.structure_to_shape_like((n, n, n), vs, (0, 0, 0), xyz, res, kind="mask"))
    assert np.array_equal(m_d, m_o) and np.abs(t_d - t_o).max() <= 1e-13 * t_o.max()
    rots = synth.random_rotations(8, seed=3)
    res_ = []
    for t, m in ((t_d, m_d), (t_o, m_o)):
        c = run_scan(pfb, case.target, t, m, rots, True)
        res_.append((c.lcc.copy(), c.rot.copy()))
    assert np.abs(res_[0][0] - res_[1][0]).max() < 1e-6 and (res_[0][1] == res_[1][1]).mean() > 0.9999


NCCL_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
rank = int(sys.argv[1])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=2,
                        device_id=torch.device("cuda", rank))
from powerfit_b200 import CUDACorrelator
g = np.load(os.path.join(%(root)r, "tests", "golden", "scan_32_plain.npz"))
target, template, mask = (g[k].astype(np.float64) for k in ("target", "template", "mask"))
c = CUDACorrelator(target, device=rank, shard=True)
c.template, c.mask, c.rotations = template, mask, g["rotations"]
c.scan()                                   # shards the rotation list, merges with one MAX all-reduce
ok = np.abs(c.lcc - g["lcc"]).max() <= 1e-4
decided = (g["lcc"] - g["lcc2"]) > 1e-4
ok = ok and np.array_equal(c.rot[decided], g["rot"][decided])
single = CUDACorrelator(target, device=rank)         # shard=False: the whole list on this rank
single.template, single.mask, single.rotations = template, mask, g["rotations"]
single.scan()
ok = ok and np.array_equal(single.lcc, c.lcc) and np.array_equal(single.rot, c.rot)
# result_rank: a MAX reduce to one rank, the other rank gets nothing (the reference's parent-process semantics)
root = CUDACorrelator(target, device=rank, shard=True, result_rank=1)
root.template, root.mask, root.rotations = template, mask, g["rotations"]
root.scan()
if rank == 1:
    ok = ok and np.array_equal(root.lcc, single.lcc) and np.array_equal(root.rot, single.rot)
else:
    ok = ok and root.lcc is None and root.rot is None
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_two_gpu_sharded_scan_equals_single_gpu(pfb, tmp_path):
    """CUDACorrelator.scan() under a 2-rank NCCL process group: rotation blocks per rank + one packed MAX
    all-reduce give, on every rank, exactly the single-GPU grids (and the reference golden)."""
    import socket, subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "nccl_worker.py"
    script.write_text(NCCL_WORKER % dict(root=root, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    codes = [p.wait(timeout=600) for p in procs]
    assert codes == [0, 0]


def test_pyramid_matches_reference_golden(pfb):
    """N4 on the device (pfb_gaussian_filter, pfb_zoom_linear behind powerfit_b200.pyramid) against maps made
    by the reference's lower_resolution / resample: the Gaussian filter bit for bit, the linear zoom to 4 ulp."""
    from powerfit_b200 import pyramid
    g = load_golden("pyramid")
    vs, res0 = float(g["voxelspacing"]), float(g["resolution"])
    levels = pyramid.image_pyramid(g["map"], vs, res0, [float(r) for r in g["targets"]], resampling_rate=2)
    for i, res in enumerate(g["targets"]):
        low = pyramid.lower_resolution(g["map"], vs, res0, float(res))
        assert np.array_equal(low, g["low_%d" % i])
        out, nvs = levels[i]
        ref = g["res_%d" % i]
        assert out.shape == ref.shape and nvs == float(g["vs_%d" % i])
        assert np.abs(out - ref).max() <= 4 * np.finfo(np.float64).eps * np.abs(ref).max()
    with pytest.raises(ValueError, match="lower than original data"):
        pyramid.image_pyramid(g["map"], vs, res0, [4.0])


def test_target_prep_matches_reference_golden(pfb):
    """N3: the CLI's resample -> trim -> extend sequence (powerfit.py:219-233) with the zoom on the device,
    against the arrays the reference functions produce."""
    from powerfit_b200 import target_prep as T
    g = load_golden("target_prep")
    arr, vs, origin = T.prepare_target(g["map"], float(g["voxelspacing"]), list(g["origin"]), float(g["resolution"]))
    ref = g["final"]
    assert arr.shape == ref.shape and vs == float(g["final_vs"]) and np.allclose(origin, g["trim_origin"], rtol=0, atol=0)
    assert np.abs(arr - ref).max() <= 4 * np.finfo(np.float64).eps * np.abs(ref).max()
    cube, _, _ = T.prepare_target(g["map"], float(g["voxelspacing"]), list(g["origin"]), float(g["resolution"]), fused=True)
    assert cube.shape == (32, 32, 32) and np.array_equal(cube[:ref.shape[0], :ref.shape[1], :ref.shape[2]], arr)


def test_map_file_to_search(pfb, tmp_path):
    """N3 end to end: a map written as .mrc is read back into page-locked memory (volume_io.read_map_f32), and the
    search on the Volume read from the file equals the search on the array that was written (float32 map)."""
    import torch
    from powerfit_b200 import synth, volume_io as V
    case = synth.make_case(n=32, voxelspacing=3.0, resolution=9.0, n_res=60, rg=8.0, n_copies=2, seed=77)
    target = case.target.astype(np.float32)
    path = str(tmp_path / "map.mrc")
    V.Volume(target, 3.0, (1.5, -3.0, 6.0)).tofile(path)
    a, vs, origin, m = V.read_map_f32(path)
    assert m._pin is not None and m._pin.is_pinned() and a.dtype == np.float32
    assert np.array_equal(a, target) and abs(vs - 3.0) < 1e-6 and list(origin) == [1.5, -3.0, 6.0]
    vol = V.Volume.fromfile(path)
    rots = synth.random_rotations(6, seed=4)
    c1 = run_scan(pfb, vol.array, case.template, case.mask, rots, True)
    c2 = run_scan(pfb, target.astype(np.float64), case.template, case.mask, rots, True)
    assert np.array_equal(c1.lcc, c2.lcc) and np.array_equal(c1.rot, c2.rot)
    V.Volume(c1.lcc, vs, origin).tofile(str(tmp_path / "lcc.mrc"))               # powerfit.py:306-308
    back, _, _ = V.parse_volume(str(tmp_path / "lcc.mrc"))
    assert np.array_equal(back.astype(np.float32), c1.lcc)
