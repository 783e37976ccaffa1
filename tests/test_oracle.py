"""CPU tests: the oracle (numpy + C restatement) against the reference's known
answers and the golden vectors produced by the real reference."""
import numpy as np
import pytest

from conftest import load_golden, golden_inputs


def test_rotate_reference_unit_vectors(oracle):
    # /root/reference/tests/test_extensions.py:10-37
    grid = np.zeros((4, 5, 6))
    for idx in [(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 0, 2), (0, 0, -1), (-1, 0, 0)]:
        grid[idx] = 1
    out = np.zeros_like(grid)
    oracle.rotate_grid3d(grid, np.eye(3), 2, out, True)
    assert np.allclose(out, grid)
    out.fill(0)
    oracle.rotate_grid3d(grid, np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]]), 2, out, False)
    answer = np.zeros_like(out)
    for idx in [(0, 0, 0), (0, 1, 0), (0, 1, -1), (0, 2, 0), (0, -1, 0), (-1, 0, 0)]:
        answer[idx] = 1
    assert np.allclose(answer, out)


def test_rotate_grids_reference_case(oracle):
    # /root/reference/tests/test_powerfitter.py:43-65 (5x6x7, radius 3)
    shape = (5, 6, 7)
    template = np.zeros(shape)
    template[0, 0, 0:3] = 1
    out = np.zeros(shape)
    oracle.rotate_grid3d(template, np.eye(3), 3, out, False)
    assert np.allclose(out, template)
    c, s = np.cos(np.radians(90)), np.sin(np.radians(90))
    out = np.zeros(shape)
    oracle.rotate_grid3d(template, np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]), 3, out, False)
    answer = np.zeros(shape)
    answer[0, :3, 0] = 1
    assert np.allclose(out, answer)


@pytest.mark.parametrize("impl", ["c", "numpy", "ref"])
def test_rotate_golden_vectors(oracle, impl):
    g = load_golden("rotate_vectors")
    if impl == "ref":
        ext = oracle.load_reference_extension()
        if ext is None:
            pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    for i in range(int(g["n"])):
        grid, R = g["grid_%d" % i], g["rotmat_%d" % i]
        radius, nearest = (int(v) for v in g["meta_%d" % i])
        out = np.zeros_like(grid)
        if impl == "c":
            oracle.rotate_grid3d(grid, R, radius, out, nearest)
        elif impl == "numpy":
            oracle.rotate_grid3d_numpy(grid, R, radius, out, nearest)
        else:
            ext.rotate_grid3d(grid, np.ascontiguousarray(R), radius, out, nearest)
        ref = g["out_%d" % i]
        if nearest:
            assert np.array_equal(out, ref), i
        else:
            assert np.allclose(out, ref, rtol=0, atol=1e-14), i
        assert np.array_equal(out != 0, ref != 0) or not nearest


def test_conj_multiply_and_calc_lcc(oracle):
    # /root/reference/tests/test__powerfit.py:13-42
    rng = np.random.default_rng(0)
    a = rng.random(100) + 1j * rng.random(100)
    b = rng.random(100) + 1j * rng.random(100)
    out = np.zeros(100, dtype=np.complex128)
    oracle.conj_multiply(a, b, out)
    assert np.allclose(out, a.conj() * b)
    gcc, ave, ave2 = rng.random(100), rng.random(100), rng.random(100) + 1
    mask = (rng.random(100) > 0.5).astype(np.uint8)
    lcc = np.zeros(100)
    oracle.calc_lcc(gcc, ave, ave2, mask, lcc)
    want = np.where(mask, gcc / np.sqrt(ave2 - ave ** 2), 0)
    assert np.allclose(lcc, want)


def test_laplace_matches_scipy(oracle):
    from scipy.ndimage import laplace
    a = np.random.default_rng(1).random((6, 7, 9))
    assert np.allclose(oracle.laplace_wrap(a), laplace(a, mode="wrap"), atol=1e-13)


def test_lcc_chain_known_answer(oracle):
    # /root/reference/tests/test_powerfitter.py:67-79
    g = load_golden("lcc_chain")
    target = g["target"]
    c = oracle.OracleCorrelator(target)
    c._lcc_mask.fill(1)
    c.template = target.copy()
    c.mask = np.ones(target.shape)
    c._rotate = lambda grid, R, r, out, nearest: out.__setitem__(slice(None), grid)
    scan = c._translational_scan(np.eye(3))
    assert abs(scan.max() - 1) < 1e-12 and scan.argmax() == 0
    assert np.allclose(scan, g["lcc_scan"], atol=1e-10)


def test_contract_errors(oracle):
    # /root/reference/tests/test_powerfitter.py:93-169
    rng = np.random.default_rng(3)
    c = oracle.OracleCorrelator(rng.random((8, 9, 10)))
    assert c._target.max() == 1
    with pytest.raises(ValueError):
        c.template = rng.random((3, 3, 3))
    with pytest.raises(ValueError):
        c.mask = np.ones((8, 9, 10))
    c.template = rng.random((8, 9, 10))
    with pytest.raises(ValueError):
        c.mask = np.zeros((8, 9, 10))
    with pytest.raises(ValueError):
        c.mask = np.ones((3, 3, 3))
    with pytest.raises(ValueError):
        c.scan()
    c.mask = np.ones((8, 9, 10))
    ind = c._mask != 0
    assert abs(c._template[ind].mean()) < 1e-12 and abs(c._template[ind].std() - 1) < 1e-12
    c.rotations = [0] * 27
    assert c._rotations.shape == (3, 3, 3)
    with pytest.raises(ValueError):
        c.rotations = [0] * 3


SMALL = ["scan_16x18x20_plain", "scan_12x15x14_laplace", "scan_24_laplace_cw", "scan_32_plain", "scan_32x64x96_cw",
         "scan_96x128x64_laplace"]


@pytest.mark.parametrize("name", SMALL)
def test_scan_matches_reference_golden(oracle, name):
    g = load_golden(name)
    target, template, mask = golden_inputs(g, name)
    c = oracle.OracleCorrelator(target, laplace=bool(g["laplace"]))
    c.template, c.mask, c.rotations = template, mask, g["rotations"]
    assert c._rmax == int(g["rmax"]) and float(c._norm_factor) == float(g["norm_factor"])
    assert np.allclose(c._template, g["prepped_template"], atol=1e-5)
    c.scan(track_second=True)
    lcc = np.nan_to_num(c.lcc, nan=0.0)
    assert np.allclose(lcc, g["lcc"], rtol=0, atol=1e-6)
    decided = (g["lcc"] - g["lcc2"]) > 1e-6
    assert np.array_equal(c.rot[decided], g["rot"][decided])
    assert np.allclose(c._lcc2[decided], g["lcc2"][decided], atol=1e-6)


def test_scan_config1_subset_matches_golden(oracle):
    """Config 1 (64^3): the first 40 rotations' running best against the golden full
    scan is not comparable, so check the oracle on a strided subset through the
    partition/merge path instead, and the full golden in the GPU suite."""
    g = load_golden("scan_config1_64")
    target, template, mask = golden_inputs(g, "scan_config1_64")
    R = g["rotations"]
    sub = R[::54]                      # 12 rotations
    lcc1, rot1 = oracle.parallel_scan(target, template, mask, sub, nproc=1)
    lcc2, rot2 = oracle.parallel_scan(target, template, mask, sub, nproc=2)
    assert np.array_equal(np.nan_to_num(lcc1), np.nan_to_num(lcc2)) and np.array_equal(rot1, rot2)
    # every voxel's best over the subset can never beat the golden best over all 648
    assert (np.nan_to_num(lcc1) <= g["lcc"] + 1e-6).all()
    hit = np.isin(g["rot"], np.arange(0, 648, 54)) & (g["lcc"] > 0)
    assert np.allclose(lcc1[hit], g["lcc"][hit], atol=1e-6)


def test_partition_matches_reference_rule(oracle):
    assert oracle.partition_rotations(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert oracle.partition_rotations(648, 8)[-1] == (567, 648)
    assert oracle.partition_rotations(5, 1) == [(0, 5)]


ANALYZER_CASES = ["scan_config1_64", "scan_32_plain", "scan_24_laplace_cw", "scan_config2_128_subset", "rough"]


def analyzer_case(name):
    """(lcc, rot, rotations, steps, voxelspacing, origin, z_sigma, positions, solutions, text)."""
    g = load_golden("analyzer_solutions")
    if name == "rough":
        lcc, rot, rotations = g["rough_lcc"], g["rough_rot"], g["rough_rotations"]
    else:
        s = load_golden(name)
        lcc, rot, rotations = s["lcc"], s["rot"], s["rotations"]
    prm = g[name + "_params"]
    return (lcc, rot, rotations, int(prm[0]), float(prm[1]), tuple(float(v) for v in prm[2:5]), float(prm[5]),
            set(tuple(int(c) for c in p) for p in g[name + "_positions"]), g[name + "_solutions"],
            str(g[name + "_text"]))


@pytest.mark.parametrize("name", ANALYZER_CASES)
def test_oracle_watershed_matches_reference_analyzer(oracle, name):
    """N1 oracle pin: feature maxima and solution rows equal the real reference Analyzer's."""
    lcc, rot, rotations, steps, vs, origin, zs, positions, solutions, _ = analyzer_case(name)
    pos = oracle.watershed_positions(lcc, steps)
    assert pos == positions
    rows = np.array(oracle.solution_rows(lcc, rotations, rot, pos, vs, origin, zs), dtype=np.float64)
    assert rows.shape == solutions.shape
    assert np.array_equal(rows[:, 0], solutions[:, 0])                 # same order, same cc
    assert np.allclose(rows, solutions, rtol=0, atol=1e-12)


SHAPE_CASES = ["small", "tiny_box", "config1"]


def shapes_case(name):
    g = load_golden("shapes")
    f = lambda k: g[name + "_" + k]
    return (tuple(int(v) for v in f("shape")), float(f("vs")), f("origin"), float(f("res")), f("xyz"), f("weights"),
            f("radii"), f("vol"), f("mask").astype(np.float64), f("mask_radii").astype(np.float64),
            f("core").astype(np.float64))


@pytest.mark.parametrize("name", SHAPE_CASES)
def test_oracle_shapes_match_reference(oracle, name):
    """N2 restatements (blur_points, dilate_points, determine_core_indices, structure_to_shape_like)
    against grids produced by the real reference (tests/golden/make_golden_shapes.py)."""
    shape, vs, origin, res, xyz, weights, radii, vol, mask, mask_r, core = shapes_case(name)
    t = oracle.structure_to_shape_like(shape, vs, origin, xyz, res, weights=weights, kind="vol")
    assert np.abs(t - vol).max() <= 1e-13 * vol.max()
    assert np.array_equal(t == 0, vol == 0)
    m = oracle.structure_to_shape_like(shape, vs, origin, xyz, res, kind="mask")
    assert np.array_equal(m, mask)
    m2 = oracle.structure_to_shape_like(shape, vs, origin, xyz, res, radii=radii.copy(), kind="mask")
    assert np.array_equal(m2, mask_r)
    assert np.array_equal(oracle.determine_core_indices(m), core)


def test_oracle_pyramid_matches_reference(oracle):
    """N4 restatements against maps produced by the real reference (tests/golden/make_golden_pyramid.py)."""
    g = load_golden("pyramid")
    vs, res0 = float(g["voxelspacing"]), float(g["resolution"])
    for i, res in enumerate(g["targets"]):
        low = oracle.lower_resolution(g["map"], vs, res0, float(res))
        assert np.array_equal(low, g["low_%d" % i])
        out, nvs = oracle.resample(low, vs, vs / (float(res) / 4.0))
        assert np.array_equal(out, g["res_%d" % i]) and nvs == float(g["vs_%d" % i])


def test_oracle_reproduces_full_search_maximum(oracle):
    """Headline-size pin (BASELINE configs[1], 128^3, Laplace, all 7416 rotations of the 10 degree set, run by
    the REAL reference: tests/golden/make_golden_full128.py).  The oracle recomputes the one rotation that wins
    at the global maximum and must reproduce the reference's LCC there; solutions.out's first row is that voxel."""
    from conftest import golden_inputs
    g = load_golden("scan_config2_128_full")
    assert g["rotations"].shape == (7416, 3, 3)
    target, template, mask = golden_inputs(g, "scan_config2_128_full")
    top = g["solutions"][0]
    steps, vs, ox, oy, oz, zs = (float(v) for v in g["analyzer_params"])
    z, y, x = (int(v) for v in g["argmax"])
    assert np.allclose(top[3:6], [x * vs + ox, y * vs + oy, z * vs + oz], rtol=0, atol=1e-9)
    assert abs(top[0] - float(g["lcc64_max"])) < 1e-12
    hits = [i for i, R in enumerate(g["rotations"]) if np.array_equal(R.ravel(), top[6:])]
    assert hits
    o = oracle.OracleCorrelator(target, laplace=True)
    o.template, o.mask, o.rotations = template, mask, g["rotations"][hits[:1]]
    o.scan()
    assert float(o._norm_factor) == float(g["norm_factor"])
    assert abs(float(o.lcc[z, y, x]) - float(g["lcc64_max"])) < 1e-9
