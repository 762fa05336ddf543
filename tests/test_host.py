"""CPU: the C-ABI library builds, loads and exports every symbol include/chan_vese_b200.h declares; the host-side
helpers (no GPU needed) agree with the oracle; compute entry points fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

import chan_vese_b200 as cv
from chan_vese_b200 import _ffi, build, synth
from oracle import coracle as co

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "chan_vese_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cvb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 50
    for n in names:
        assert hasattr(handle, n), "missing export: " + n
    assert set(names) == set(_ffi.SIGNATURES), "ctypes table and header disagree"
    assert b"sm_100a" in _ffi.lib().cvb_version()


def test_library_contains_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    """Without a device every compute entry point must fail (CVB_ERR_NO_DEVICE), never compute on the CPU."""
    if _ffi.lib().cvb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cv.ChanVeseError) as e:
        cv.Context(0)
    assert e.value.status == _ffi.ERR_NO_DEVICE
    with pytest.raises(cv.ChanVeseError):
        cv.perona_malik([np.zeros((4, 4), np.uint8)], 4, 4, 10.0, 0.25, 1.0)
    src = open(os.path.join(ROOT, "chan_vese_b200", "solver.py")).read() + open(os.path.join(ROOT, "bench.py")).read()
    assert "import oracle" not in open(os.path.join(ROOT, "chan_vese_b200", "solver.py")).read()
    for f in os.listdir(os.path.join(ROOT, "chan_vese_b200")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "chan_vese_b200", f)).read().replace("the oracle", ""), f


def test_pm_num_steps_matches_fp_loop():
    for L, T in [(0.25, 100.0), (0.1, 1.5), (0.25, 20.0), (0.2, 0.7), (0.1, 0.35), (0.25, 0.0), (0.05, 1.0), (0.3, 0.3)]:
        assert cv.pm_num_steps(L, T) == synth.pm_steps_expected(L, T) == co.oracle().cvo_pm_num_steps(L, T)
    assert cv.pm_num_steps(0.25, 100.0) == 400 and cv.pm_num_steps(0.1, 1.5) == 15
    assert cv.pm_num_steps(0.0, 1.0) == -1


def test_level_set_initialisers(kat):
    assert np.array_equal(cv.levelset_checkerboard(23, 37), kat["checker_23x37"])
    assert np.array_equal(cv.levelset_rect(20, 30, 4, 5, 11, 7), kat["rect_20x30"])
    assert np.array_equal(cv.levelset_circ(40, 50, 22, 19, 13), kat["circ_40x50"])
    assert np.array_equal(cv.levelset_circ(30, 30, 3, 27, 9), kat["circ_clip"])
    for h, w in [(1, 1), (5, 6), (250, 370), (64, 1001)]:
        assert np.array_equal(cv.levelset_checkerboard(h, w), co.levelset_checkerboard(h, w))
    assert np.array_equal(cv.levelset_rect(10, 10, -3, 7, 6, 9), co.levelset_rect(10, 10, -3, 7, 6, 9))
    with pytest.raises(cv.ChanVeseError):
        cv.levelset_checkerboard(0, 5)


def test_slab_partition_covers_image_and_aligns_to_groups():
    for h, rows in [(16384, 32), (4096, 32), (1000, 8), (430, 4), (250, 4)]:
        for nranks in (1, 2, 4, 8):
            spans = [cv.slab_partition(h, rows, nranks, r) for r in range(nranks)]
            assert spans[0][0] == 0 and spans[-1][1] == h
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a < b and c < d
            assert all(lo % rows == 0 for lo, _ in spans)
    assert cv.slab_partition(16384, 32, 8, 3) == (6144, 8192)
    with pytest.raises(cv.ChanVeseError):
        [cv.slab_partition(100, 32, 8, r) for r in range(8)]  # fewer tiles than ranks: some slab would be empty
    with pytest.raises(cv.ChanVeseError):
        cv.slab_partition(1000, 8, 3, 0)  # the rank count must divide the 32 reduction groups
    assert 48 <= cv.auto_tile_rows(16384, 16384) <= 192 and 48 <= cv.auto_tile_rows(16384, 16384, 1, 8) <= 192
    assert cv.auto_tile_rows(430, 640) == 4


def test_synthetic_inputs_are_deterministic():
    a, b = synth.seastar(), synth.seastar()
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[0].shape == (250, 370)
    whole = synth.hashed_scene_rows(512, 384, 0, 512, cell=128)
    part = synth.hashed_scene_rows(512, 384, 200, 330, cell=128)
    assert all(np.array_equal(w[200:330], p) for w, p in zip(whole, part))  # slabs agree with the whole image
    assert np.array_equal(synth.batch_images(3, 2, 64, 64)[1], np.stack(synth.batch_image(4, 64, 64)))


def test_tile_choice_does_not_depend_on_the_rank_count_and_balances_the_slabs():
    """The tiling fixes the order of the fused sums, so the automatic tile length must be the same for every GPU count
    (bit-identical results at 1, 2, 4, 8 GPUs); it is chosen so that the slabs of 2, 4 and 8 ranks are balanced."""
    for h, w in [(16384, 16384), (8192, 8192), (12000, 9000), (4096, 4096), (430, 640)]:
        t = cv.auto_tile_rows(h, w, 1, 1)
        assert all(cv.auto_tile_rows(h, w, 1, n) == t for n in (2, 4, 8, 16))
    h = w = 16384
    t = cv.auto_tile_rows(h, w)
    for n in (2, 4, 8):
        rows = [hi - lo for lo, hi in (cv.slab_partition(h, t, n, r) for r in range(n))]
        assert sum(rows) == h and max(rows) <= 1.01 * h / n, (n, rows)


def test_release_scratch_and_trim_are_declared_and_epsilon_is_validated():
    from chan_vese_b200 import frontend as fe
    with pytest.raises(fe.MsgExit) as e:
        fe._validate(fe._parser().parse_args(["-i", __file__, "-e", "0"]))
    assert "smoothing parameter" in str(e.value)
