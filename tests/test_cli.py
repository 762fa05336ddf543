"""The C++14 host front-end (cli/chan_vese.cpp -> bin/chan_vese): the reference's option surface and error behaviour
(src/main.cpp:756-874) on CPU; the end-to-end run against the Python binding on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from chan_vese_b200 import build, synth


@pytest.fixture(scope="module")
def cli():
    return build.build_cli()


def _run(cli, *args):
    return subprocess.run([cli] + list(args), capture_output=True, text=True)


def write_ppm(path, planes):  # planes B,G,R
    h, w = planes[0].shape
    rgb = np.stack([planes[2], planes[1], planes[0]], axis=-1)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(rgb.tobytes())


def read_pnm(path):
    with open(path, "rb") as f:
        magic = f.readline().strip()
        w, h = map(int, f.readline().split())
        assert f.readline().strip() == b"255"
        data = np.frombuffer(f.read(), dtype=np.uint8)
    if magic == b"P5":
        return [data.reshape(h, w)]
    rgb = data.reshape(h, w, 3)
    return [rgb[..., 2], rgb[..., 1], rgb[..., 0]]  # back to B,G,R


def test_option_validation_messages(cli, tmp_path):
    img = tmp_path / "x.ppm"
    write_ppm(img, synth.seastar(20, 24))
    cases = [
        ([], "Error: you have to specify input file name!"),
        (["-i", str(tmp_path / "nope.ppm")], "does not exists!"),
        (["-i", str(img), "--dt", "0"], "Cannot have negative or zero timestep"),
        (["-i", str(img), "--mu", "-1"], "Length penalty parameter cannot be negative"),
        (["-i", str(img), "--lambda1", "1", "2"], "Number of lambda1 values must be 3 for a colored input image."),
        (["-i", str(img), "-g", "--lambda2", "1", "2"], "Too many lambda2 values for a grayscale image."),
        (["-i", str(img), "--lambda1", "1", "-2", "1"], "Any value of lambda1 cannot be negative."),
        (["-i", str(img), "-L", "0.3"], "must be between 0 and 0.25"),
        (["-i", str(img), "-L", "0.25", "-T", "0.1"], "The segmentation duration must exceed"),
        (["-i", str(img), "-P", "XX"], "Invalid text position requested."),
        (["-i", str(img), "-l", "pink"], "Invalid contour color requested."),
        (["-i", str(img), "-R", "-C"], "Cannot initialize with both rectangular and circular contour"),
        (["-i", str(img), "--frobnicate"], "error: unrecognised option '--frobnicate'"),
        (["-i", str(img), "--dt"], "error: the required argument for option '--dt' is missing"),
    ]
    for args, msg in cases:
        r = _run(cli, *args)
        assert r.returncode == 1 and msg in r.stderr and r.stdout == "", (args, r.stderr)
    r = _run(cli, "--help")
    assert r.returncode == 0 and "--max-steps" in r.stdout and "-S [ --segment ]" in r.stdout


@pytest.mark.gpu
def test_cli_end_to_end_matches_binding(cli, tmp_path, ctx):
    """README.md:53-style run (reduced PM time) through the binary: files named like the reference's, silent stdout,
    contents equal to the library called through the Python binding."""
    import chan_vese_b200 as cv
    planes = synth.seastar(120, 150, seed=3)
    img = tmp_path / "star.ppm"
    write_ppm(img, planes)
    r = _run(cli, "-i", str(img), "-s", "-N", "30", "-S", "-L", "0.25", "-T", "5", "-K", "30", "--nu", "-0.5", "--lambda1", "1", "1", "0.5",
             "-V", "--stats")
    assert r.returncode == 0 and r.stdout == "", r.stderr
    assert "steps=" in r.stderr
    ref = ctx.segment(planes, cv.levelset_checkerboard(120, 150), cv.make_params(nu=-0.5, lambda1=[1, 1, 0.5]), tol=1e-3,
                      max_steps=30, smooth=True, K=30.0, L=0.25, T=5.0)
    pm = read_pnm(tmp_path / "star_pm.ppm")
    assert all(np.array_equal(a, b) for a, b in zip(pm, ref["pm"]))
    sel = read_pnm(tmp_path / "star_selection.ppm")
    m = ref["mask"].astype(bool)
    for k in range(3):
        assert np.array_equal(sel[k][m], planes[k][m]) and np.all(sel[k][~m] == 255)
    assert os.path.exists(tmp_path / "star_contour.ppm")
    assert "steps=%d " % ref["steps"] in r.stderr
    # -V: the per-step frame stream (PPM images back to back): frame 0 + one frame per step (src/main.cpp:929, :997), fed by
    # the asynchronous mask observer; the last frame shows the contour of the final level set
    raw = open(tmp_path / "star.ppms", "rb").read()
    head = b"P6\n150 120\n255\n"
    fsz = len(head) + 120 * 150 * 3
    assert len(raw) == (1 + ref["steps"]) * fsz and all(raw[k * fsz:k * fsz + len(head)] == head for k in range(1 + ref["steps"]))
    assert raw[-fsz:] == open(tmp_path / "star_contour.ppm", "rb").read()
    first = np.frombuffer(raw[len(head):fsz], dtype=np.uint8).reshape(120, 150, 3)
    assert (first != np.stack([planes[2], planes[1], planes[0]], axis=-1)).any()  # the checkerboard's contour is drawn
    # grayscale + non-interactive circle init + inverted selection
    r = _run(cli, "-i", str(img), "-g", "-s", "-I", "-N", "10", "--circ", "75,60,30")
    assert r.returncode == 0, r.stderr
    assert read_pnm(tmp_path / "star_selection.ppm")[0].shape == (120, 150)
