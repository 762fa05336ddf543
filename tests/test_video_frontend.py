"""The Python front-end (chan_vese_b200/frontend.py): the reference's command line with cv2 image I/O and the per-step
XVID video (src/main.cpp:583-1008, src/VideoWriterManager.cpp).  CPU tests drive the host logic with the ORACLE as the
compute backend (test infrastructure: the product backend is the CUDA library and has no CPU path); the GPU test runs
the real thing."""
import os

import cv2
import numpy as np
import pytest

import chan_vese_b200 as cv
from chan_vese_b200 import frontend as fe
from chan_vese_b200 import synth
from oracle import coracle as co


class OracleBackend:
    """perona_malik / csv_run / mask with the signatures of chan_vese_b200.Context, computed by the C oracle."""

    def perona_malik(self, channels, K, L, T):
        return co.perona_malik(channels, K, L, T)

    def csv_run(self, channels, u, p, tol=1e-3, max_steps=-1, frame=None):
        op = co.params(p.mu, p.nu, p.dt, p.eps, list(p.lambda1), list(p.lambda2))
        stop = co.stop_condition(channels, tol)
        steps, nrm = 0, float("nan")
        limit = max_steps if max_steps >= 0 else 2**31 - 1
        while steps < limit:  # src/main.cpp:963-1001
            u, nrm, _, _ = co.csv_step(channels, u, op)
            steps += 1
            if frame is not None and frame(u, steps):
                break
            if nrm <= stop:
                break
        return u, steps, nrm

    def mask(self, u, invert=False):
        return co.mask(u, invert)


def write_png(path, planes):  # planes B,G,R
    cv2.imwrite(str(path), np.stack(planes, axis=-1))


def run(args, backend=None, **kw):
    try:
        return fe.run([str(a) for a in args], backend=backend, **kw), ""
    except fe.MsgExit as e:
        return 1, str(e)


def test_add_suffix_and_saturate():
    assert fe.add_suffix("/a/b/star.png", "pm") == "/a/b/star_pm.png"  # src/main.cpp:158-167
    assert fe.add_suffix("star", "selection") == "star_selection"
    u = np.array([[-3.0, 0.4, 0.5, 0.50001, 1.5, 2.5, 300.0]])
    assert fe.saturate_u8(u).tolist() == [[0, 0, 0, 1, 2, 2, 255]]  # half to even, clamped


def test_validation_messages(tmp_path):
    img = tmp_path / "x.png"
    write_png(img, synth.seastar(20, 24))
    cases = [
        ([], "Error: you have to specify input file name!"),
        (["-i", tmp_path / "nope.png"], "does not exists!"),
        (["-i", img, "--dt", "0"], "Cannot have negative or zero timestep"),
        (["-i", img, "--mu", "-1"], "Length penalty parameter cannot be negative"),
        (["-i", img, "--lambda1", "1", "2"], "Number of lambda1 values must be 3 for a colored input image."),
        (["-i", img, "-g", "--lambda2", "1", "2"], "Too many lambda2 values for a grayscale image."),
        (["-i", img, "--lambda1", "1", "-2", "1"], "Any value of lambda1 cannot be negative."),
        (["-i", img, "-L", "0.3"], "must be between 0 and 0.25"),
        (["-i", img, "-L", "0.25", "-T", "0.1"], "The segmentation duration must exceed"),
        (["-i", img, "-P", "XX"], "Invalid text position requested."),
        (["-i", img, "-l", "pink"], "Invalid contour color requested."),
        (["-i", img, "--rect", "1,2,3,4", "--circ", "5,5,2"], "Cannot initialize with both rectangular and circular contour"),
        (["-i", img, "-R"], "give the rectangle as --rect"),
        (["-i", img, "--circ", "5,5,0"], "non-zero dimensions"),
        (["-i", img, "--frobnicate"], "error: "),
    ]
    for args, msg in cases:
        rc, err = run(args, backend=OracleBackend())
        assert rc == 1 and msg in err, (args, err)
    bad = tmp_path / "notimage.png"
    bad.write_bytes(b"not an image")
    rc, err = run(["-i", bad], backend=OracleBackend())
    assert rc == 1 and "probably not an image" in err


def test_video_writer_manager_frames(tmp_path):
    """draw_contour's threshold is saturate_cast<uchar>(u) > 0 (src/VideoWriterManager.cpp:65-66), the contour is
    drawn in the requested colour on a copy of the ORIGINAL image, the overlay text colour follows the patch under it."""
    h, w = 60, 80
    img = np.full((h, w, 3), 200, np.uint8)
    u = np.full((h, w), -1.0)
    u[20:40, 30:50] = 5.0
    u[5:8, 5:8] = 0.4  # positive, but rounds to 0: no contour there
    vwm = fe.VideoWriterManager(str(tmp_path / "in.png"), img, fe.COLORS["red"], 10, "TopLeft", True)
    assert vwm.filename.endswith("in.avi")
    frame = vwm.compose(u, "t = 3")
    red = np.all(frame == np.array(fe.COLORS["red"], np.uint8), axis=-1)
    assert red[20, 30:50].all() and red[39, 30:50].all() and red[20:40, 30].all() and red[20:40, 49].all()
    assert not red[21:39, 31:49].any() and not red[5:8, 5:8].any()
    assert np.array_equal(img, np.full((h, w, 3), 200, np.uint8))  # the underlying image is never drawn on
    # bright patch -> black text (255 - 200 < 105); dark image -> white text
    color, p = vwm.overlay_color("t = 3")
    assert color == fe.COLORS["black"] and p[0] == 5
    dark = fe.VideoWriterManager(str(tmp_path / "dark.png"), np.zeros((h, w, 3), np.uint8), fe.COLORS["blue"], 10, "BottomRight", True)
    color, p = dark.overlay_color("t = 3")
    assert color == fe.COLORS["white"] and p[1] == h - 5
    for k in range(4):
        vwm.write_frame(u, "t = %d" % k)
    vwm.release()
    dark.release()
    cap = cv2.VideoCapture(vwm.filename)
    n = 0
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        assert fr.shape == (h, w, 3)
        n += 1
    assert n == 4 == vwm.frames
    # a level set without any contour must not crash (the reference would index an empty hierarchy)
    assert vwm.compose(np.full((h, w), -1.0)).shape == (h, w, 3)


def test_run_writes_the_reference_outputs(tmp_path):
    """-S -s -V -O on a small colour PNG: "_pm" and "_selection" next to the input, "<stem>.avi" with 1 + steps
    frames, contents equal to the same pipeline called directly."""
    h, w = 48, 64
    planes = synth.seastar(h, w, seed=7)
    img = tmp_path / "star.png"
    write_png(img, planes)
    be = OracleBackend()
    rc, err = run(["-i", img, "-S", "-L", "0.25", "-T", "1", "-K", "30", "-N", "6", "-s", "-V", "-O", "-l", "green", "--nu", "-0.5"],
                  backend=be)
    assert rc == 0, err
    pm_ref, n_pm = co.perona_malik(planes, 30.0, 0.25, 1.0)
    assert n_pm == 4
    pm = cv2.imread(str(tmp_path / "star_pm.png"), cv2.IMREAD_COLOR)
    assert all(np.array_equal(pm[..., k], pm_ref[k]) for k in range(3))
    u_ref, steps, _ = be.csv_run(pm_ref, cv.levelset_checkerboard(h, w), cv.make_params(nu=-0.5), 1e-3, 6)
    sel = cv2.imread(str(tmp_path / "star_selection.png"), cv2.IMREAD_COLOR)
    m = co.mask(u_ref).astype(bool)
    orig = np.stack(planes, axis=-1)
    assert np.array_equal(sel[m], orig[m]) and np.all(sel[~m] == 255)  # original pixels, not the smoothed ones (:1005)
    cap = cv2.VideoCapture(str(tmp_path / "star.avi"))
    n = 0
    while cap.read()[0]:
        n += 1
    assert n == 1 + steps  # t = 0 and one frame per step (:929, :997)


def test_run_grayscale_circle_inverted(tmp_path):
    h, w = 40, 56
    planes = synth.seastar(h, w, seed=9)
    img = tmp_path / "g.png"
    write_png(img, planes)
    frames = []

    class Sink:
        def write(self, f):
            frames.append(f.copy())

    rc, err = run(["-i", img, "-g", "--circ", "28,20,10", "-N", "3", "-s", "-I", "-V"], backend=OracleBackend(), video_writer=Sink())
    assert rc == 0, err
    gray = cv2.imread(str(img), cv2.IMREAD_GRAYSCALE)
    u_ref, steps, _ = OracleBackend().csv_run([gray], cv.levelset_circ(h, w, 28, 20, 10), cv.make_params(nch=1), 1e-3, 3)
    sel = cv2.imread(str(tmp_path / "g_selection.png"), cv2.IMREAD_COLOR)
    m = co.mask(u_ref, True).astype(bool)
    assert np.array_equal(sel[m][:, 0], gray[m]) and np.all(sel[~m] == 255)
    assert len(frames) == 1 + steps and frames[0].shape == (h, w, 3)
    # frame 0 shows the initial ring in the default colour (blue), src/main.cpp:929
    blue = np.all(frames[0] == np.array(fe.COLORS["blue"], np.uint8), axis=-1)
    assert blue.any()


def test_no_cpu_fallback(tmp_path):
    """Without a backend the front-end opens the CUDA library; on a host without a GPU that must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    img = tmp_path / "x.png"
    write_png(img, synth.seastar(20, 24))
    with pytest.raises(cv.ChanVeseError):
        fe.run(["-i", str(img), "-N", "1"])


@pytest.mark.gpu
def test_frontend_on_the_gpu(tmp_path, ctx):
    h, w = 96, 128
    planes = synth.seastar(h, w, seed=11)
    img = tmp_path / "star.png"
    write_png(img, planes)
    assert fe.main(["-i", str(img), "-S", "-L", "0.25", "-T", "2", "-K", "30", "-N", "12", "-s", "-V", "-O"]) == 0
    ref = ctx.segment(planes, cv.levelset_checkerboard(h, w), cv.make_params(), tol=1e-3, max_steps=12, smooth=True, K=30.0,
                      L=0.25, T=2.0)
    pm = cv2.imread(str(tmp_path / "star_pm.png"), cv2.IMREAD_COLOR)
    assert all(np.array_equal(pm[..., k], ref["pm"][k]) for k in range(3))
    sel = cv2.imread(str(tmp_path / "star_selection.png"), cv2.IMREAD_COLOR)
    m = ref["mask"].astype(bool)
    orig = np.stack(planes, axis=-1)
    assert np.array_equal(sel[m], orig[m]) and np.all(sel[~m] == 255)
    cap = cv2.VideoCapture(str(tmp_path / "star.avi"))
    n = 0
    while cap.read()[0]:
        n += 1
    assert n == 1 + ref["steps"]
