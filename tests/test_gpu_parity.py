"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle and the golden fixtures.

Tolerances (BASELINE.json north_star: "level-set function in fp64 within a stated relative tolerance ... final
segmentation mask with >= 99.9 % pixel agreement"):
  * strict math, one step with given region means: BIT-EXACT level set (the arithmetic is the reference's);
  * fast math (production): rel-L2(u) <= 1e-6 after full runs (measured ~1e-9), identical step
    counts, mask agreement 100 % on the fixtures (>= 99.9 % required);
  * PM uint8 planes: bit-exact in strict mode; fast mode >= 99.99 % equal and |delta| <= 1 (ties at x.5).
"""
import numpy as np
import pytest

import chan_vese_b200 as cv
from chan_vese_b200 import synth
from conftest import SMALL, rel_l2
from oracle import coracle as co

pytestmark = pytest.mark.gpu

TOL_U = 1e-6


def _p(kat, name, n, mk):
    v = kat[name + "_params"]
    return mk(v[0], v[1], v[2], v[3], list(v[4:4 + n]), list(v[7:7 + n]), nch=n)


def _planes_close(a, b, frac=0.9999):
    a, b = np.stack(a).astype(np.int16), np.stack(b).astype(np.int16)
    assert np.abs(a - b).max() <= 1
    assert (a == b).mean() >= frac


@pytest.fixture()
def strict(ctx):
    ctx.set_math_mode(True)
    yield ctx
    ctx.set_math_mode(False)


# ---- single kernels ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SMALL)
def test_curvature(ctx, kat, name):
    u = kat[name + "_u"]
    k = ctx.curvature(u)
    np.testing.assert_allclose(k, kat[name + "_kappa"], rtol=0, atol=2e-14)
    ctx.set_math_mode(True)
    try:
        assert np.array_equal(ctx.curvature(u), kat[name + "_kappa"])  # bit-exact with cv::filter2D arithmetic
    finally:
        ctx.set_math_mode(False)


@pytest.mark.parametrize("name", SMALL)
def test_region_means_stop_mask(ctx, kat, name):
    img = list(kat[name + "_img"])
    u = kat[name + "_u"]
    c1, c2 = ctx.region_means(img, u, 1.0)
    np.testing.assert_allclose(c1, kat[name + "_c1"], rtol=1e-12)
    np.testing.assert_allclose(c2, kat[name + "_c2"], rtol=1e-12)
    assert abs(ctx.stop_condition(img, 1e-3) - kat[name + "_stop"]) <= 1e-13 * kat[name + "_stop"]
    assert np.array_equal(ctx.mask(u), kat[name + "_mask"])
    assert np.array_equal(ctx.mask(u, True), 1 - kat[name + "_mask"])
    h, w = u.shape
    assert cv.region_variance(img[0], u, h, w, cv.Region.Inside, ctx=ctx) == pytest.approx(kat[name + "_c1"][0], rel=1e-12)
    assert cv.region_variance(img[0], u, h, w, cv.Region.Outside, ctx=ctx) == pytest.approx(kat[name + "_c2"][0], rel=1e-12)


def test_delta_map_parallel_pixel_function(ctx, kat):
    xs = np.ascontiguousarray(kat["hd_x"])
    d = xs.copy()
    ctx.delta_map(d, 0.7)
    assert np.array_equal(d, kat["delta"])  # IEEE div/mul/add in the reference's order
    m = xs[:2200].reshape(40, 55).copy()
    cv.ParallelPixelFunction(m, 55, 0.7, ctx=ctx)(0, m.size)
    assert np.array_equal(m.ravel(), kat["delta"][:2200])
    e = np.zeros(0)
    ctx.delta_map(e, 1.0)  # empty range is a no-op


@pytest.mark.parametrize("name", SMALL)
def test_csv_step_strict_bit_exact(strict, kat, name):
    """One step with the oracle's region means given: every other operation is the reference's, in its order."""
    img = list(kat[name + "_img"])
    n = len(img)
    u = kat[name + "_u"]
    h, w = u.shape
    c1, c2 = kat[name + "_c1"], kat[name + "_c2"]
    ref_u1, ref_norm, _, _ = co.csv_step(img, u, _p(kat, name, n, co.params), c1, c2)
    with cv.Session(strict, n, h, w) as s:
        s.upload_image(img)
        s.upload_levelset(u)
        nrm = s.csv_step(_p(kat, name, n, cv.make_params), c1, c2)
        u1 = s.download_levelset()
    assert np.array_equal(u1, ref_u1)
    assert abs(nrm - ref_norm) <= 1e-12 * ref_norm
    assert rel_l2(u1, kat[name + "_u1"]) < 1e-13  # and the cv2 result (its own c1/c2 summation order)


@pytest.mark.parametrize("name", SMALL)
def test_csv_run_small(ctx, kat, name):
    img = list(kat[name + "_img"])
    n = len(img)
    u, steps, nrm = ctx.csv_run(img, kat[name + "_u"], _p(kat, name, n, cv.make_params), tol=0.0, max_steps=5)
    assert steps == 5
    assert rel_l2(u, kat[name + "_u5"]) < 1e-10
    assert abs(nrm - kat[name + "_norm5"]) <= 1e-9 * abs(kat[name + "_norm5"])


@pytest.mark.parametrize("name", SMALL)
def test_pm_small(ctx, kat, name):
    img = list(kat[name + "_img"])
    out, n = ctx.perona_malik(img, 12.0, 0.2, 0.7)
    assert n == int(kat[name + "_pmsteps"])
    _planes_close(out, kat[name + "_pm"], frac=0.999)
    ctx.set_math_mode(True)
    try:
        out, _ = ctx.perona_malik(img, 12.0, 0.2, 0.7)
        assert np.array_equal(np.stack(out), kat[name + "_pm"])
    finally:
        ctx.set_math_mode(False)


def test_pm_step_counts_and_zero_steps(ctx):
    rng = np.random.default_rng(3)
    img = [rng.integers(0, 256, size=(33, 47), dtype=np.uint8)]
    for L, T in [(0.25, 0.25), (0.25, 0.5), (0.2, 0.7), (0.1, 0.35)]:
        out, n = ctx.perona_malik(img, 15.0, L, T)
        ref, nr = co.perona_malik(img, 15.0, L, T)
        assert n == nr == synth.pm_steps_expected(L, T)
        _planes_close(out, ref, frac=0.999)
    out, n = ctx.perona_malik(img, 15.0, 0.25, 0.0)
    assert n == 0 and np.array_equal(out[0], img[0])
    flat = [np.full((20, 300), 77, dtype=np.uint8)]
    out, _ = ctx.perona_malik(flat, 10.0, 0.25, 5.0)
    assert np.array_equal(out[0], flat[0])  # a constant image is a fixed point


# ---- the BASELINE configurations ---------------------------------------------------------------------------------
def test_config1(ctx, golden_c1):
    """README.md:53: PM -L 0.25 -T 100 -K 30 (400 steps) + CSV -N 70, checkerboard init."""
    c = synth.CONFIGS["C1"]
    img = synth.seastar()
    u0 = cv.levelset_checkerboard(c["h"], c["w"])
    assert np.array_equal(u0, co.levelset_checkerboard(c["h"], c["w"]))
    r = ctx.segment(img, u0, cv.make_params(), tol=1e-3, max_steps=70, smooth=True, **c["pm"])
    _planes_close(r["pm"], golden_c1["pm"])
    # CSV on the oracle's PM planes so that a 1-LSB PM tie does not leak into the level-set comparison
    u, steps, nrm = ctx.csv_run(list(golden_c1["pm"]), u0, cv.make_params(), tol=1e-3, max_steps=70)
    assert steps == int(golden_c1["steps"])
    assert rel_l2(u, golden_c1["u"]) < TOL_U
    assert abs(nrm - float(golden_c1["norm"])) <= 1e-6 * float(golden_c1["norm"])
    ref_mask = np.unpackbits(golden_c1["mask"])[:c["h"] * c["w"]].reshape(c["h"], c["w"])
    assert (ctx.mask(u) == ref_mask).mean() == 1.0
    assert (r["mask"] == ref_mask).mean() >= 0.999  # end to end, including the PM hand-off in uint8


def test_config2(ctx, golden_c2):
    """README.md:59-62: --dt 0.001 -t 1e-6 --nu -293 --lambda1 1 1 0.1, PM -L 0.1 -T 1.5 -K 1000 (15 steps), -N 132."""
    c = synth.CONFIGS["C2"]
    k = c["csv"]
    img = synth.night_lights()
    out, n = ctx.perona_malik(img, **c["pm"])
    assert n == 15 == int(golden_c2["pmsteps"])
    _planes_close(out, golden_c2["pm"])
    p = cv.make_params(nu=k["nu"], dt=k["dt"], lambda1=k["lambda1"])
    u, steps, nrm = ctx.csv_run(list(golden_c2["pm"]), cv.levelset_checkerboard(c["h"], c["w"]), p, tol=k["tol"],
                                max_steps=k["max_steps"])
    assert steps == int(golden_c2["steps"])
    assert rel_l2(u[::4, ::4], golden_c2["u_sub"]) < TOL_U
    assert abs(np.linalg.norm(u) - float(golden_c2["u_norm"])) <= TOL_U * float(golden_c2["u_norm"])
    ref_mask = np.unpackbits(golden_c2["mask"])[:c["h"] * c["w"]].reshape(c["h"], c["w"])
    assert (ctx.mask(u) == ref_mask).mean() >= 0.999
    # the whole level set (the fixture keeps every 4th sample only): the C oracle on the same planes, all 132 steps
    ref, rs, rn = co.csv_run(list(golden_c2["pm"]), co.levelset_checkerboard(c["h"], c["w"]),
                             co.params(nu=k["nu"], dt=k["dt"], lambda1=k["lambda1"]), k["tol"], k["max_steps"])
    assert rs == steps and rel_l2(u, ref) < TOL_U and abs(nrm - rn) <= 1e-6 * rn
    assert np.array_equal(ctx.mask(u), co.mask(ref))


def test_config3_reduced_ring_init(ctx):
    """C3 (grayscale, one-pixel ring init of InteractiveDataCirc) at 512^2 for 40 steps against the oracle."""
    img = synth.two_phase(512, 512, seed=3, discs=10)
    u0 = cv.levelset_circ(512, 512, 256, 256, 128)
    assert np.array_equal(u0, co.levelset_circ(512, 512, 256, 256, 128))
    u, steps, _ = ctx.csv_run(img, u0, cv.make_params(nch=1), tol=0.0, max_steps=40)
    ref, rs, _ = co.csv_run(img, u0, co.params(), 0.0, 40)
    assert steps == rs == 40
    assert rel_l2(u, ref) < TOL_U
    assert (ctx.mask(u) == co.mask(ref)).mean() >= 0.999


def test_config3_all_2000_steps_reduced_size(ctx):
    """C3's step count -- 2000 iterations, tolerance 0, grayscale, ring init -- at 256^2 (the oracle needs ~10 s): the
    level set stays within the tolerance over the whole run, not just over its first steps."""
    n = 256
    img = synth.two_phase(n, n, seed=3, discs=6)
    u0 = cv.levelset_circ(n, n, n // 2, n // 2, n // 4)
    u, steps, nrm = ctx.csv_run(img, u0, cv.make_params(nch=1), tol=0.0, max_steps=2000)
    ref, rs, rn = co.csv_run(img, u0, co.params(), 0.0, 2000)
    assert steps == rs == 2000
    assert rel_l2(u, ref) < TOL_U
    assert abs(nrm - rn) <= 1e-6 * max(rn, 1e-300)
    assert (ctx.mask(u) == co.mask(ref)).mean() >= 0.999


def test_early_stop_same_step_as_oracle(ctx):
    """The tolerance ends the run (not -N): same breaking step, and its update is applied (src/main.cpp:994,1000)."""
    img = synth.seastar(120, 150, seed=11)
    u0 = cv.levelset_checkerboard(120, 150)
    for tol in (0.05, 0.2):
        ref, rs, rn = co.csv_run(img, u0, co.params(), tol, 200)
        u, steps, nrm = ctx.csv_run(img, u0, cv.make_params(), tol=tol, max_steps=200)
        assert rs < 200 and steps == rs
        assert rel_l2(u, ref) < TOL_U
        assert abs(nrm - rn) <= 1e-6 * rn


def test_unlimited_steps_and_frame_observer(ctx):
    img = synth.seastar(64, 80, seed=5)
    u0 = cv.levelset_checkerboard(64, 80)
    ref, rs, _ = co.csv_run(img, u0, co.params(), 0.1, 10 ** 6)
    seen = []

    def frame(u, step):
        seen.append((step, float(u.sum())))
        return 0

    u, steps, _ = ctx.csv_run(img, u0, cv.make_params(), tol=0.1, max_steps=-1, frame=frame)  # -N -1 -> unlimited
    assert steps == rs
    assert [s for s, _ in seen] == list(range(1, steps + 1))
    assert seen[-1][1] == pytest.approx(float(u.sum()), rel=1e-12)
    with pytest.raises(cv.ChanVeseError) as e:
        ctx.csv_run(img, u0, cv.make_params(), tol=0.0, max_steps=5, frame=lambda u, s: s == 2)
    assert e.value.status == 7  # CVB_ERR_CALLBACK


# ---- resident sessions, batches, decomposition invariance -------------------------------------------------------------
def test_session_matches_one_shot_and_device_checkerboard(ctx):
    img = synth.seastar(90, 131, seed=2)
    p = cv.make_params(lambda1=[1.0, 0.5, 2.0])
    u_ref, s_ref, _ = ctx.csv_run(img, cv.levelset_checkerboard(90, 131), p, tol=0.0, max_steps=12)
    with cv.Session(ctx, 3, 90, 131) as s:
        s.upload_image(img)
        s.init_checkerboard()
        assert np.array_equal(s.download_levelset(), cv.levelset_checkerboard(90, 131))
        steps, _ = s.csv_run(p, tol=0.0, max_steps=12)
        assert steps == s_ref and np.array_equal(s.download_levelset(), u_ref)
        # continuing a run from the resident level set == running longer from the start
        s.csv_run(p, tol=0.0, max_steps=3)
        u15 = s.download_levelset()
    u_long, _, _ = ctx.csv_run(img, cv.levelset_checkerboard(90, 131), p, tol=0.0, max_steps=15)
    assert np.array_equal(u15, u_long)


def test_tile_rows_invariance(ctx):
    """Results must not depend on the tiling beyond summation order (1e-12), and are bit-reproducible run to run."""
    img = synth.seastar(200, 300, seed=9)
    u0 = cv.levelset_checkerboard(200, 300)
    res = {}
    try:
        for rows in (4, 8, 32, 64):
            ctx.set_tile_rows(rows)
            res[rows] = ctx.csv_run(img, u0, cv.make_params(), tol=0.0, max_steps=25)[0]
        again = ctx.csv_run(img, u0, cv.make_params(), tol=0.0, max_steps=25)[0]
    finally:
        ctx.set_tile_rows(0)
    assert np.array_equal(again, res[64])
    for rows in (4, 8, 32):
        assert rel_l2(res[rows], res[64]) < 1e-11
    pm = {}
    try:
        for rows in (4, 32):
            ctx.set_tile_rows(rows)
            pm[rows] = ctx.perona_malik(img, 30.0, 0.25, 3.0)[0]
    finally:
        ctx.set_tile_rows(0)
    assert all(np.array_equal(a, b) for a, b in zip(pm[4], pm[32]))  # PM has no reductions: bit-identical


def test_batch_equals_sessions(ctx):
    """Image-parallel batches (BASELINE config 5): independent c1/c2, stop flags and step counts per image."""
    count, h, w = 6, 96, 112
    imgs = synth.batch_images(0, count, h, w)
    u0 = cv.levelset_checkerboard(h, w)
    p = cv.make_params()
    ctx.set_tile_rows(8)  # the automatic tiling depends on the image count; fix it so the sums add in the same order
    try:
        with cv.Batch(ctx, count, 3, h, w) as b:
            b.upload_images(imgs)
            b.upload_levelset(u0)
            assert b.perona_malik(30.0, 0.25, 2.0) == 8
            steps, norms = b.csv_run(p, tol=0.08, max_steps=60)
            got = [(b.download_image(m), b.download_levelset(m), b.mask(m)) for m in range(count)]
        single = [ctx.segment(list(imgs[m]), u0, p, tol=0.08, max_steps=60, smooth=True, K=30.0, L=0.25, T=2.0)
                  for m in range(count)]
    finally:
        ctx.set_tile_rows(0)
    assert len(set(steps.tolist())) > 1  # early stops differ per image
    for m in range(count):
        r = single[m]
        assert all(np.array_equal(a, b) for a, b in zip(got[m][0], r["pm"]))
        assert steps[m] == r["steps"]
        assert np.array_equal(got[m][1], r["u"])
        assert np.array_equal(got[m][2], r["mask"])
        assert norms[m] == r["norm"]
    # and against the oracle for one of them
    pm_ref, _ = co.perona_malik(list(imgs[2]), 30.0, 0.25, 2.0)
    _planes_close(got[2][0], pm_ref, frac=0.999)
    u_ref, s_ref, _ = co.csv_run(got[2][0], u0, co.params(), 0.08, 60)
    assert steps[2] == s_ref and rel_l2(got[2][1], u_ref) < TOL_U


# ---- full BASELINE sizes ---------------------------------------------------------------------------------------------
def test_config3_full_size_few_steps(ctx):
    """4096^2 grayscale, ring init: 3 steps against the oracle (a few seconds of CPU)."""
    img = synth.two_phase()
    u0 = cv.levelset_circ(4096, 4096, 2048, 2048, 1024)
    u, steps, nrm = ctx.csv_run(img, u0, cv.make_params(nch=1), tol=0.0, max_steps=3)
    ref, rs, rn = co.csv_run(img, u0, co.params(), 0.0, 3)
    assert steps == rs == 3
    assert rel_l2(u, ref) < 1e-9
    assert abs(nrm - rn) <= 1e-9 * rn
    assert (ctx.mask(u) == co.mask(ref)).mean() >= 0.999


def test_config4_full_size_windows(ctx):
    """16384^2 RGB: one PM step and one CSV step with given region means are local stencils, so windows of the
    full-size result must equal the oracle run on the same window plus its halo -- bit-exact in strict math."""
    h = w = 16384
    rng = np.random.default_rng(4)
    img = [rng.integers(0, 256, size=(h, w), dtype=np.uint8) for _ in range(3)]
    c1, c2 = np.array([100.0, 120.5, 90.25]), np.array([140.0, 99.0, 131.0])
    p_gpu, p_ref = cv.make_params(lambda1=[1.0, 0.5, 2.0]), co.params(lambda1=[1.0, 0.5, 2.0])
    ctx.set_math_mode(True)
    try:
        with cv.Session(ctx, 3, h, w) as s:
            s.upload_image(img)
            s.init_checkerboard()
            assert s.perona_malik(25.0, 0.25, 0.5) == 2
            pm = s.download_image()
            s.csv_step(p_gpu, c1, c2)
            u1 = s.download_levelset()
            m1 = s.mask()
    finally:
        ctx.set_math_mode(False)
    u0 = None
    wins = [(0, 0), (0, w - 96), (h - 96, 0), (h - 96, w - 96), (8000, 8100), (4090, 12345), (16000, 300)]
    for (r0, c0) in wins:
        ra, rb = max(r0 - 8, 0), min(r0 + 96 + 8, h)
        ca, cb = max(c0 - 8, 0), min(c0 + 96 + 8, w)
        # PM window: two steps reach 4 pixels; compare the inner 96x96 (or up to the true border)
        sub = [np.ascontiguousarray(x[ra:rb, ca:cb]) for x in img]
        ref_pm = [co.pm_evolve(x, 25.0, 0.25, 2) for x in sub]
        ir = slice(0 if ra == 0 else 8, (rb - ra) if rb == h else (rb - ra - 8))
        ic = slice(0 if ca == 0 else 8, (cb - ca) if cb == w else (cb - ca - 8))
        for k in range(3):
            q = np.clip(np.rint(ref_pm[k]), 0, 255).astype(np.uint8)
            # windows that do not touch the true border see clamped neighbours at their cut edges: inner part only
            assert np.array_equal(q[ir, ic], pm[k][ra:rb, ca:cb][ir, ic])
        if u0 is None:
            u0 = cv.levelset_checkerboard(h, w)
        subpm = [np.ascontiguousarray(x[ra:rb, ca:cb]) for x in pm]
        ref_u1, _, _, _ = co.csv_step(subpm, u0[ra:rb, ca:cb], p_ref, c1, c2)
        assert np.array_equal(ref_u1[ir, ic], u1[ra:rb, ca:cb][ir, ic])
    assert np.array_equal(m1, (u1.astype(np.float32) > 0).astype(np.uint8))


@pytest.fixture(scope="module")
def bench_scene():
    """The bench's own 16384^2 RGB input (synth.hashed_scene_rows, BASELINE configs[3])."""
    import os
    return synth.hashed_scene_rows(16384, 16384, 0, 16384, threads=min(8, os.cpu_count() or 1))


@pytest.mark.parametrize("tile_rows", [0, 186, 81])
def test_config4_full_size_production_kernels(ctx, bench_scene, tile_rows):
    """The kernels bench.py TIMES, at the geometry it times them: 16384^2 RGB, fast math (cp.async row rings), the
    automatic tile length and the two of round 1 (186: one GPU, 81: eight).  PM 20 steps, then 6 CSV steps with given
    region means -- both are local stencils then, so windows of the full-size result are compared with the oracle run on
    the window plus a halo wider than the stencils reach (PM 20 steps: 40 pixels; CSV 6 steps: 12).  Windows sit on all
    four image borders and corners, across tile seams (rows) and strip seams (columns) and in the interior.
    Tolerances: PM planes |delta| <= 1 and >= 99.99 % equal (ties at x.5); level set rel-L2 <= 1e-9 and max-abs 1e-9 x
    scale per window; mask identical."""
    h = w = 16384
    K, L, T, npm, ncsv = 10.0, 0.25, 5.0, 20, 6
    c1, c2 = np.array([171.3, 139.9, 101.2]), np.array([93.7, 108.4, 125.6])
    p_gpu, p_ref = cv.make_params(lambda1=[1.0, 0.5, 2.0]), co.params(lambda1=[1.0, 0.5, 2.0])
    ctx.set_tile_rows(tile_rows)
    try:
        tr = tile_rows or cv.auto_tile_rows(h, w, 1, 1)
        with cv.Session(ctx, 3, h, w) as s:
            s.upload_image(bench_scene)
            s.init_checkerboard()
            assert s.perona_malik(K, L, T) == npm
            pm = s.download_image()
            for _ in range(ncsv):
                s.csv_step(p_gpu, c1, c2)
            u = s.download_levelset()
            m = s.mask()
    finally:
        ctx.set_tile_rows(0)
    u0 = cv.levelset_checkerboard(h, w)
    seam = (h // 2 // tr) * tr          # a tile seam near the middle
    strip = (w // 2 // 62) * 62         # a CSV strip seam; PM strips are 60 wide
    pstrip = (w // 3 // 60) * 60
    wins = [(0, 0), (0, w - 96), (h - 96, 0), (h - 96, w - 96), (0, 7000), (h - 96, 9000), (5000, 0), (11000, w - 96),
            (seam - 48, strip - 48), (seam - 48, pstrip - 48), (tr - 40, 300), (h - tr - 50, 12000), (8111, 8222)]
    HP, HC = 48, 16
    worst_rel, pm_diff, pm_tot = 0.0, 0, 0
    for (r0, c0) in wins:
        # ---- PM: oracle on the raw scene window + 48
        ra, rb, ca, cb = max(r0 - HP, 0), min(r0 + 96 + HP, h), max(c0 - HP, 0), min(c0 + 96 + HP, w)
        ir = slice(r0 - ra, r0 - ra + 96)
        ic = slice(c0 - ca, c0 - ca + 96)
        for k in range(3):
            ref = co.pm_evolve(np.ascontiguousarray(bench_scene[k][ra:rb, ca:cb]), K, L, npm)
            q = np.clip(np.rint(ref), 0, 255).astype(np.int16)[ir, ic]
            d = np.abs(q - pm[k][r0:r0 + 96, c0:c0 + 96].astype(np.int16))
            assert d.max() <= 1, (tile_rows, r0, c0, k)
            pm_diff += int((d != 0).sum())
            pm_tot += d.size
        # ---- CSV: oracle on the GPU's own PM planes (so a PM tie cannot leak into the level-set comparison) + 16
        ra, rb, ca, cb = max(r0 - HC, 0), min(r0 + 96 + HC, h), max(c0 - HC, 0), min(c0 + 96 + HC, w)
        ir = slice(r0 - ra, r0 - ra + 96)
        ic = slice(c0 - ca, c0 - ca + 96)
        sub = [np.ascontiguousarray(x[ra:rb, ca:cb]) for x in pm]
        ur = np.ascontiguousarray(u0[ra:rb, ca:cb])
        for _ in range(ncsv):
            ur, _, _, _ = co.csv_step(sub, ur, p_ref, c1, c2)
        a, b = u[r0:r0 + 96, c0:c0 + 96], ur[ir, ic]
        rel = rel_l2(a, b)
        worst_rel = max(worst_rel, rel)
        assert rel <= 1e-9, (tile_rows, r0, c0, rel)
        assert np.abs(a - b).max() <= 1e-9 * max(np.abs(b).max(), 1.0), (tile_rows, r0, c0)
        assert np.array_equal(m[r0:r0 + 96, c0:c0 + 96], co.mask(b)), (tile_rows, r0, c0)
    assert pm_diff <= 1e-4 * pm_tot, (pm_diff, pm_tot)
    print("tile_rows=%d (%d): worst window rel-L2 %.2e, PM pixels off by one LSB %d of %d" % (tile_rows, tr, worst_rel, pm_diff, pm_tot))


# ---- fp32 variant (reported separately: judged on the mask) ----------------------------------------------------------------
def test_fp32_variant_config1(ctx, golden_c1):
    """CVB_PRECISION_F32 on BASELINE config 1: PM planes within 1 LSB, mask agreement >= 99.9 % with the fp64 reference
    result; the level set itself is only reported (fp32 cannot meet the fp64 tolerance)."""
    c = synth.CONFIGS["C1"]
    h, w = c["h"], c["w"]
    img = synth.seastar()
    ref_mask = np.unpackbits(golden_c1["mask"])[:h * w].reshape(h, w)
    with cv.Session(ctx, 3, h, w, fp32=True) as s:
        s.upload_image(img)
        assert s.perona_malik(**c["pm"]) == 400
        pm = s.download_image()
        d = np.abs(np.stack(pm).astype(int) - golden_c1["pm"].astype(int))
        assert d.max() <= 1 and (d == 0).mean() >= 0.99
        s.upload_image(list(golden_c1["pm"]))
        s.init_checkerboard()
        steps, _ = s.csv_run(cv.make_params(), tol=1e-3, max_steps=70)
        u = s.download_levelset()
        m = s.mask()
    assert steps == int(golden_c1["steps"])
    assert (m == ref_mask).mean() >= 0.999
    print("fp32 C1: rel-L2(u) = %.3e, mask agreement = %.5f" % (rel_l2(u, golden_c1["u"]), (m == ref_mask).mean()))


def test_fp32_variant_gray_ring_and_batch(ctx):
    img = synth.two_phase(384, 448, seed=3, discs=8)
    u0 = cv.levelset_circ(384, 448, 224, 192, 96)
    ref, rs, _ = co.csv_run(img, u0, co.params(), 0.0, 40)
    with cv.Session(ctx, 1, 384, 448, fp32=True) as s:
        s.upload_image(img)
        s.upload_levelset(u0)
        steps, _ = s.csv_run(cv.make_params(nch=1), tol=0.0, max_steps=40)
        m = s.mask()
        u = s.download_levelset()
    assert steps == rs == 40
    assert (m == co.mask(ref)).mean() >= 0.999
    assert rel_l2(u, ref) < 1e-2
    imgs = synth.batch_images(0, 3, 96, 112)
    with cv.Batch(ctx, 3, 3, 96, 112, fp32=True) as b:
        b.upload_images(imgs)
        b.upload_levelset(cv.levelset_checkerboard(96, 112))
        b.perona_malik(30.0, 0.25, 1.0)
        steps, _ = b.csv_run(cv.make_params(), tol=0.0, max_steps=15)
        masks = [b.mask(k) for k in range(3)]
    for k in range(3):
        r = ctx.segment(list(imgs[k]), cv.levelset_checkerboard(96, 112), cv.make_params(), tol=0.0, max_steps=15, smooth=True,
                        K=30.0, L=0.25, T=1.0)
        assert (masks[k] == r["mask"]).mean() >= 0.995


def test_overlapped_upload_and_packed_mask(ctx):
    """cvb_session_upload_image_smooth == upload_image + perona_malik (bit-identical); the packed mask is
    numpy.packbits of the byte mask."""
    img = synth.seastar(130, 203, seed=8)
    with cv.Session(ctx, 3, 130, 203) as s:
        s.upload_image(img)
        n1 = s.perona_malik(25.0, 0.25, 2.0)
        ref = s.download_image()
        n2 = s.upload_image_smooth(img, 25.0, 0.25, 2.0)
        got = s.download_image()
        assert n1 == n2 == 8 and all(np.array_equal(a, b) for a, b in zip(ref, got))
        s.init_checkerboard()
        s.csv_run(cv.make_params(), tol=0.0, max_steps=10)
        m = s.mask()
        assert np.array_equal(s.mask_packed(), np.packbits(m, axis=1))
        assert np.array_equal(s.mask_packed(invert=True), np.packbits(1 - m, axis=1))


def test_error_paths_and_degenerate_shapes(ctx):
    """Invalid arguments come back as status codes with a message (never a crash, never a CPU result); the smallest
    shapes the reference accepts (1x1, 1xN, Nx1) run."""
    import ctypes as C
    from chan_vese_b200 import _ffi
    img = [np.zeros((8, 8), np.uint8)] * 2
    with pytest.raises(cv.ChanVeseError) as e:
        ctx.perona_malik(img, 10.0, 0.25, 1.0)  # two channels: the reference knows 1 (-g) or 3
    assert e.value.status == _ffi.ERR_INVALID_ARGUMENT and "n must be 1 or 3" in str(e.value)
    with pytest.raises(cv.ChanVeseError):
        ctx.csv_run([np.zeros((8, 8), np.uint8)], np.zeros((8, 8)), cv.make_params(eps=0.0, nch=1), max_steps=1)
    with pytest.raises(cv.ChanVeseError):
        ctx.perona_malik([np.zeros((8, 8), np.uint8)], 10.0, 0.0, 1.0)  # L must be > 0
    lib = _ffi.lib()
    assert lib.cvb_csv_run(ctx._h, None, 1, 8, 8, None, None, 1e-3, 1, None, None, C.cast(None, _ffi.FRAME_FN), None) != _ffi.OK
    hd = C.c_void_p()
    assert lib.cvb_session_create(ctx._h, 1, 0, 8, 0, C.byref(hd)) == _ffi.ERR_INVALID_ARGUMENT
    assert lib.cvb_session_create_slab(ctx._h, 1, 64, 64, 3, 40, 0, C.byref(hd)) == _ffi.ERR_INVALID_ARGUMENT  # unaligned slab
    assert b"aligned" in lib.cvb_last_error(ctx._h) or b"split" in lib.cvb_last_error(ctx._h)
    with pytest.raises(ValueError):
        cv.Session(ctx, 1, 8, 8).upload_levelset(np.zeros((4, 4)))
    for h, w in [(1, 1), (1, 7), (9, 1), (2, 3)]:
        rng = np.random.default_rng(h * 10 + w)
        im = [rng.integers(0, 256, size=(h, w), dtype=np.uint8)]
        u0 = rng.standard_normal((h, w))
        # tol < 0 never stops (the reference lets it through); with tol = 0 a 1x1 image sits on the knife
        # edge du = 0 <= 0, which rounding decides
        u, steps, _ = ctx.csv_run(im, u0, cv.make_params(nch=1), tol=-1.0, max_steps=4)
        ref, rs, _ = co.csv_run(im, u0, co.params(), -1.0, 4)
        assert steps == rs == 4 and rel_l2(u, ref) < 1e-10
        out, n = ctx.perona_malik(im, 10.0, 0.25, 1.0)
        refp, _ = co.perona_malik(im, 10.0, 0.25, 1.0)
        assert n == 4 and np.abs(out[0].astype(int) - refp[0].astype(int)).max() <= 1
    # max_steps = 0: nothing runs, u comes back unchanged
    u0 = np.arange(64.0).reshape(8, 8)
    u, steps, _ = ctx.csv_run([np.zeros((8, 8), np.uint8)], u0, cv.make_params(nch=1), tol=0.0, max_steps=0)
    assert steps == 0 and np.array_equal(u, u0)


def test_random_shapes_against_oracle(ctx):
    """Widths around the 62- and 60-column strip boundaries, odd sizes, heights around the tile lengths, random
    parameters: CSV (3-6 steps) and PM (2-5 steps) against the oracle."""
    rng = np.random.default_rng(2024)
    widths = [2, 3, 59, 60, 61, 62, 63, 64, 65, 119, 120, 121, 123, 124, 125, 126, 185, 186, 187, 247, 248, 249, 250, 311, 373]
    heights = [2, 3, 4, 5, 7, 8, 9, 31, 33, 47, 48, 49, 50, 97, 130]
    for case in range(30):
        w = int(widths[case % len(widths)])
        h = int(rng.choice(heights))
        n = int(rng.choice([1, 3]))
        img = [rng.integers(0, 256, size=(h, w), dtype=np.uint8) for _ in range(n)]
        u0 = rng.standard_normal((h, w)) * float(rng.choice([0.05, 1.0, 30.0]))
        lam1 = [float(x) for x in rng.uniform(0.2, 2.0, n)] if case % 2 else [1.0] * n
        lam2 = [float(x) for x in rng.uniform(0.2, 2.0, n)] if case % 2 else [1.0] * n
        kw = dict(mu=float(rng.uniform(0, 1)), nu=float(rng.uniform(-1, 1)), dt=float(rng.choice([0.1, 1.0])),
                  eps=float(rng.choice([0.5, 1.0, 2.0])), lambda1=lam1, lambda2=lam2, nch=n)
        steps = int(rng.integers(3, 7))
        u, s, nrm = ctx.csv_run(img, u0, cv.make_params(**kw), tol=-1.0, max_steps=steps)
        ref, rs, rn = co.csv_run(img, u0, co.params(**kw), -1.0, steps)
        assert s == rs == steps, (case, h, w, n)
        assert rel_l2(u, ref) < 1e-9, (case, h, w, n, rel_l2(u, ref))
        assert abs(nrm - rn) <= 1e-9 * max(rn, 1e-300), (case, h, w, n)
        K, L = float(rng.choice([5.0, 30.0])), float(rng.choice([0.1, 0.25]))
        T = L * int(rng.integers(2, 6)) - L / 2
        out, npm = ctx.perona_malik(img, K, L, T)
        refp, nr = co.perona_malik(img, K, L, T)
        assert npm == nr
        d = np.abs(np.stack(out).astype(int) - np.stack(refp).astype(int))
        assert d.max() <= 1 and (d == 0).mean() >= 0.995, (case, h, w, n, d.max(), (d == 0).mean())


# ---- Perona-Malik temporal blocking: two steps per launch must equal two launches, bit for bit --------------------
def _pm_both_ways(ctx, img, K, L, T, batch=None):
    import os
    out = []
    for fuse in ("0", "1"):
        os.environ["CVB_PM_FUSE"] = fuse
        try:
            if batch is None:
                h, w = img[0].shape
                with cv.Session(ctx, len(img), h, w) as s:
                    s.upload_image(img)
                    n = s.perona_malik(K, L, T)
                    planes = s.download_image() + (s.download_pm_state() if n >= 2 else [])  # uint8 result + fp64 state
            else:
                with cv.Batch(ctx, *batch) as b:
                    b.upload_images(img)
                    n = b.perona_malik(K, L, T)
                    planes = [p for m in range(batch[0]) for p in b.download_image(m)]
        finally:
            os.environ.pop("CVB_PM_FUSE", None)
        out.append((planes, n))
    return out


@pytest.mark.parametrize("shape", [(1, 70), (2, 9), (3, 1), (5, 57), (11, 56), (64, 113), (97, 300), (250, 370), (513, 1025)])
@pytest.mark.parametrize("nsteps", [3, 4, 5, 8, 11])
def test_pm_two_steps_per_launch_bit_identical(ctx, shape, nsteps):
    """pm2_step_kernel (temporal blocking) against one launch per step: identical uint8 planes for every shape class
    (one-row / one-column images, strips that touch both borders, ragged last strips, several segments) and for step
    counts that leave zero or one unfused fp64 step."""
    h, w = shape
    rng = np.random.default_rng(1000 * h + w + nsteps)
    img = [rng.integers(0, 256, size=(h, w), dtype=np.uint8) for _ in range(3)]
    L = 0.25
    T = L * (nsteps - 0.5)
    (a, na), (b, nb) = _pm_both_ways(ctx, img, 12.0, L, T)
    assert na == nb == nsteps
    for k in range(6):  # 3 uint8 planes and the 3 fp64 planes before the last step: bit for bit
        assert np.array_equal(a[k], b[k]), (shape, nsteps, k, int((a[k] != b[k]).sum()))
    # and both agree with the oracle within the PM tolerance
    ref, nr = co.perona_malik(img, 12.0, L, T)
    assert nr == nsteps
    _planes_close(b[:3], ref, frac=0.999 if h * w < 2000 else 0.9999)


def test_pm_two_steps_per_launch_batch_and_large(ctx):
    rng = np.random.default_rng(77)
    imgs = rng.integers(0, 256, size=(5, 3, 96, 130), dtype=np.uint8)
    (a, na), (b, nb) = _pm_both_ways(ctx, imgs, 20.0, 0.2, 1.3, batch=(5, 3, 96, 130))
    assert na == nb == 7
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    img = synth.hashed_scene_rows(4096, 4096, 0, 4096, threads=4)
    (a, na), (b, nb) = _pm_both_ways(ctx, img, 10.0, 0.25, 5.0)
    assert na == nb == 20
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


# ---- asynchronous per-step mask observer (src/main.cpp:997, VideoWriterManager.cpp:57-75) -----------------
@pytest.mark.parametrize("contour_rule", [True, False])
@pytest.mark.parametrize("tol,max_steps", [(0.0, 23), (0.05, 200)])
def test_async_mask_observer_matches_the_level_set_frames(ctx, contour_rule, tol, max_steps):
    """The masks that travel through the pinned ring while later steps run are, step for step, the threshold of the
    level sets the synchronous observer shows -- same steps, same order, the breaking step included, nothing after it."""
    h, w = 91, 150
    img = synth.seastar(h, w, seed=31)
    u0 = cv.levelset_checkerboard(h, w)
    prm = cv.make_params()
    frames, masks = [], []
    u1, s1, n1 = ctx.csv_run(img, u0, prm, tol=tol, max_steps=max_steps, frame=lambda u, step: frames.append((step, u.copy())) and 0)
    u2, s2, n2 = ctx.csv_run_masks(img, u0, prm, lambda m, step: masks.append((step, m.copy())) and 0, tol=tol,
                                   max_steps=max_steps, contour_rule=contour_rule)
    assert s1 == s2 and n1 == n2 and np.array_equal(u1, u2)
    assert [s for s, _ in frames] == [s for s, _ in masks] == list(range(1, s1 + 1))
    if tol > 0:
        assert s1 < max_steps  # the tolerance ended the run
    for (_, u), (_, m) in zip(frames, masks):
        ref = (np.clip(np.rint(u), 0, 255) > 0) if contour_rule else (u.astype(np.float32) > 0)
        assert np.array_equal(m.astype(bool), ref)


def test_async_mask_observer_abort(ctx):
    img = synth.seastar(64, 80, seed=3)
    u0 = cv.levelset_checkerboard(64, 80)
    seen = []
    with pytest.raises(cv.ChanVeseError) as e:
        ctx.csv_run_masks(img, u0, cv.make_params(), lambda m, step: seen.append(step) or step == 3, tol=0.0, max_steps=50)
    assert "CALLBACK" in str(e.value) and seen == [1, 2, 3]


# ---- BASELINE configs[4]: the image batch --------------------------------------------------------------------------------
def test_config5_batch_subsample_against_oracle(ctx):
    """64 of the 4096 images of C5 (512 x 512 RGB, PM 40 steps, CSV <= 50 steps with per-image early stop) as ONE batch
    job against the oracle image by image: identical step counts, rel-L2(u) <= 1e-6, masks identical, PM planes within
    one LSB.  The oracle runs the images on a thread pool (it is plain C behind ctypes: the GIL is released)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    c = synth.CONFIGS["C5"]
    h, w, n, count = c["h"], c["w"], c["n"], 64
    imgs = synth.batch_images(0, count, h, w)
    prm = cv.make_params()
    with cv.Batch(ctx, count, n, h, w) as b:
        b.upload_images(imgs)
        b.init_checkerboard()
        assert b.perona_malik(**c["pm"]) == 40
        steps, norms = b.csv_run(prm, tol=1e-3, max_steps=c["csv"]["max_steps"])
        packed = b.masks_packed()
        sample = {m: (b.download_image(m), b.download_levelset(m)) for m in range(count)}
    u0 = co.levelset_checkerboard(h, w)

    def ref(m):
        pm, npm = co.perona_malik(list(imgs[m]), c["pm"]["K"], c["pm"]["L"], c["pm"]["T"])
        u, s, nrm = co.csv_run(pm, u0, co.params(), 1e-3, c["csv"]["max_steps"])
        return pm, u, s, nrm

    with ThreadPoolExecutor(min(16, os.cpu_count() or 1)) as ex:
        refs = list(ex.map(ref, range(count)))
    worst = 0.0
    for m, (pm_ref, u_ref, s_ref, n_ref) in enumerate(refs):
        pm_gpu, u_gpu = sample[m]
        _planes_close(pm_gpu, pm_ref, frac=0.999)
        # the CSV comparison must not inherit a PM tie (one LSB of one pixel): when the planes differ, the oracle is re-run on
        # the GPU's own planes
        if any(not np.array_equal(a, r) for a, r in zip(pm_gpu, pm_ref)):
            u_ref, s_ref, n_ref = co.csv_run(pm_gpu, u0, co.params(), 1e-3, c["csv"]["max_steps"])
        assert steps[m] == s_ref, (m, steps[m], s_ref)
        r = rel_l2(u_gpu, u_ref)
        worst = max(worst, r)
        assert r <= TOL_U, (m, r)
        mask = np.unpackbits(packed[m], axis=1)[:, :w]
        assert np.array_equal(mask, co.mask(u_ref)), m
        assert abs(norms[m] - n_ref) <= 1e-6 * max(n_ref, 1e-300)
    assert steps.min() >= 1 and steps.max() <= c["csv"]["max_steps"]
    print("C5 subsample: steps %d..%d, worst rel-L2 %.2e" % (steps.min(), steps.max(), worst))


# ---- the tiling of a job: segments fix the order of the sums, segments per CTA must not matter ---------------------
def test_results_do_not_depend_on_segments_per_cta(ctx):
    """Geom::seg_mult (how many segments one CTA marches through -- chosen per GPU count for speed) changes neither the
    level set, nor the step at which the tolerance ends the run, nor the norm: bit for bit."""
    import os
    h, w = 203, 330
    img = synth.seastar(h, w, seed=9)
    u0 = cv.levelset_checkerboard(h, w)
    prm = cv.make_params(lambda1=[1.0, 0.7, 1.3])
    for tile, tol, max_steps in ((8, 0.0, 17), (8, 0.3, 60), (12, 0.3, 60), (40, 0.0, 9)):
        out = []
        ctx.set_tile_rows(tile)
        try:
            for mult in ("1", "2", "3", "7"):
                os.environ["CVB_SEG_MULT"] = mult
                with cv.Session(ctx, 3, h, w) as s:
                    s.upload_image(img)
                    s.upload_levelset(u0)
                    steps, norm = s.csv_run(prm, tol=tol, max_steps=max_steps)
                    out.append((s.download_levelset(), steps, norm))
        finally:
            os.environ.pop("CVB_SEG_MULT", None)
            ctx.set_tile_rows(0)
        for u, steps, norm in out[1:]:
            assert steps == out[0][1] and norm == out[0][2] and np.array_equal(u, out[0][0]), (tile, tol)
        ref, rs, _ = co.csv_run(img, u0, co.params(lambda1=[1.0, 0.7, 1.3]), tol, max_steps)
        assert rs == out[0][1] and rel_l2(out[0][0], ref) < TOL_U


def test_prefetch_image_streams_images_through_a_session(ctx):
    """prefetch_image / restore_image: image k+1 travels to the device while image k is being solved; every image gets
    exactly the result of a plain upload."""
    h, w = 120, 170
    imgs = [synth.seastar(h, w, seed=s, arms=5 + s) for s in (41, 42, 43)]
    prm = cv.make_params()

    def solve(s):
        s.init_checkerboard()
        n = s.perona_malik(20.0, 0.25, 1.5)
        steps, norm = s.csv_run(prm, tol=0.0, max_steps=9)
        return s.download_image(), s.download_levelset(), n, steps, norm

    ref = []
    for im in imgs:
        with cv.Session(ctx, 3, h, w) as s:
            s.upload_image(im)
            ref.append(solve(s))
    with cv.Session(ctx, 3, h, w) as s:
        s.prefetch_image(imgs[0])
        for k in range(3):
            s.restore_image()
            if k + 1 < 3:
                s.prefetch_image(imgs[k + 1])  # runs behind the solve below
            got = solve(s)
            assert got[2:] == ref[k][2:]
            assert all(np.array_equal(a, b) for a, b in zip(got[0], ref[k][0])) and np.array_equal(got[1], ref[k][1])
