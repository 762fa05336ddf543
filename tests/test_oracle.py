"""CPU: the C oracle (oracle/cv_oracle.c) against the golden vectors produced by real OpenCV arithmetic
(oracle/cv2_oracle.py via oracle/make_golden.py), and the pass-structured port (ref_cpu.cpp) against the oracle."""
import numpy as np
import pytest

from conftest import SMALL, rel_l2
from chan_vese_b200 import synth
from oracle import coracle as co


def _params(kat, name, n):
    v = kat[name + "_params"]
    return co.params(v[0], v[1], v[2], v[3], list(v[4:4 + n]), list(v[7:7 + n]), nch=n)


@pytest.mark.parametrize("name", SMALL)
def test_curvature_bit_exact(kat, name):
    assert np.array_equal(co.curvature(kat[name + "_u"]), kat[name + "_kappa"])


@pytest.mark.parametrize("name", SMALL)
def test_csv_step_matches_cv2(kat, name):
    img = list(kat[name + "_img"])
    p = _params(kat, name, len(img))
    u1, nrm, c1, c2 = co.csv_step(img, kat[name + "_u"], p)
    # the only non-bit-exact ingredient is numpy's vs C's evaluation order of (1 + 2/pi*atan)/2 sums
    np.testing.assert_allclose(c1, kat[name + "_c1"], rtol=1e-14)
    np.testing.assert_allclose(c2, kat[name + "_c2"], rtol=1e-14)
    assert rel_l2(u1, kat[name + "_u1"]) < 1e-14
    assert abs(nrm - kat[name + "_norm1"]) <= 1e-12 * abs(kat[name + "_norm1"])
    u5, steps, nrm5 = co.csv_run(img, kat[name + "_u"], p, 0.0, 5)
    assert steps == 5
    assert rel_l2(u5, kat[name + "_u5"]) < 1e-12


@pytest.mark.parametrize("name", SMALL)
def test_pm_stop_mask(kat, name):
    img = list(kat[name + "_img"])
    pm, n = co.perona_malik(img, 12.0, 0.2, 0.7)
    assert n == int(kat[name + "_pmsteps"]) == synth.pm_steps_expected(0.2, 0.7)
    assert np.array_equal(np.stack(pm), kat[name + "_pm"])
    assert abs(co.stop_condition(img, 1e-3) - kat[name + "_stop"]) <= 1e-13 * kat[name + "_stop"]
    assert np.array_equal(co.mask(kat[name + "_u"]), kat[name + "_mask"])
    assert np.array_equal(co.mask(kat[name + "_u"], True), 1 - kat[name + "_mask"])


def test_inits_and_scalars(kat):
    assert np.array_equal(co.levelset_checkerboard(23, 37), kat["checker_23x37"])
    assert np.array_equal(co.levelset_rect(20, 30, 4, 5, 11, 7), kat["rect_20x30"])
    assert np.array_equal(co.levelset_circ(40, 50, 22, 19, 13), kat["circ_40x50"])
    assert np.array_equal(co.levelset_circ(30, 30, 3, 27, 9), kat["circ_clip"])
    o = co.oracle()
    xs = kat["hd_x"]
    assert np.array_equal(np.array([o.cvo_heaviside(x, 0.7) for x in xs]), kat["heaviside"])
    assert np.array_equal(np.array([o.cvo_delta(x, 0.7) for x in xs]), kat["delta"])
    assert np.array_equal(co.delta_map(xs, 0.7), kat["delta"])
    assert co.oracle().cvo_pm_num_steps(0.25, 100.0) == 400
    assert co.oracle().cvo_pm_num_steps(0.1, 1.5) == 15
    assert co.oracle().cvo_pm_num_steps(0.25, 20.0) == 80


def test_config1_full(golden_c1):
    """BASELINE config 1 (README.md:53) end to end: PM 400 steps + CSV 70 steps."""
    c = synth.CONFIGS["C1"]
    pm, n = co.perona_malik(synth.seastar(), **c["pm"])
    assert n == 400 == int(golden_c1["pmsteps"])
    assert np.array_equal(np.stack(pm), golden_c1["pm"])
    u, steps, nrm = co.csv_run(pm, co.levelset_checkerboard(c["h"], c["w"]), co.params(), 1e-3, 70)
    assert steps == int(golden_c1["steps"])
    assert rel_l2(u, golden_c1["u"]) < 1e-12
    assert np.array_equal(np.packbits(co.mask(u)), golden_c1["mask"])


def test_refcpu_port_equals_oracle():
    """The timed CPU baseline (reference pass/thread structure) computes what the oracle computes."""
    rng = np.random.default_rng(7)
    img = [rng.integers(0, 256, size=(40, 52), dtype=np.uint8) for _ in range(3)]
    a, na = co.perona_malik(img, 20.0, 0.25, 2.0, impl="oracle")
    b, nb = co.perona_malik(img, 20.0, 0.25, 2.0, impl="refcpu")
    assert na == nb == 8
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    p = co.params(lambda1=[1.0, 0.5, 2.0])
    u0 = co.levelset_checkerboard(40, 52)
    ua, sa, _ = co.csv_run(img, u0, p, 1e-3, 12, impl="oracle")
    ub, sb, _ = co.csv_run(img, u0, p, 1e-3, 12, impl="refcpu")
    assert sa == sb
    assert rel_l2(ub, ua) < 1e-13
