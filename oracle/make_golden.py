"""ORACLE tooling: generate tests/golden/*.npz by running the cv2 transcription (oracle/cv2_oracle.py), i.e.
real OpenCV arithmetic on the reference's call sites, on the deterministic synthetic inputs of
chan_vese_b200/synth.py.  Run once in the build container (cv2 4.13 present); the fixtures travel to the GPU box.

    python -m oracle.make_golden
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chan_vese_b200 import synth  # noqa: E402
from oracle import cv2_oracle as cvo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def small_cases():
    rng = np.random.default_rng(1234)
    d = {}
    for name, (h, w, n) in {"rgb": (24, 31, 3), "gray": (17, 40, 1), "thin": (1, 9, 1), "tall": (11, 1, 1),
                            "two": (2, 2, 3)}.items():
        ch = [rng.integers(0, 256, size=(h, w), dtype=np.uint8) for _ in range(n)]
        u = rng.standard_normal((h, w)) * rng.choice([0.01, 1.0, 50.0], size=(h, w))
        lam1 = [1.0, 0.7, 1.3][:n]
        lam2 = [0.9, 1.1, 1.0][:n]
        mu, nu, dt, eps = 0.5, 0.1, 0.5, 1.0
        d[name + "_img"] = np.stack(ch)
        d[name + "_u"] = u
        d[name + "_kappa"] = cvo.curvature(u)
        u1, nrm, c1, c2 = cvo.csv_step(ch, u, mu, nu, dt, eps, lam1, lam2)
        d[name + "_u1"], d[name + "_norm1"], d[name + "_c1"], d[name + "_c2"] = u1, nrm, np.array(c1), np.array(c2)
        u5, steps, nrm5 = cvo.csv_run(ch, u, mu, nu, dt, eps, lam1, lam2, 0.0, 5)
        d[name + "_u5"], d[name + "_norm5"] = u5, nrm5
        d[name + "_stop"] = cvo.stop_condition(ch, 1e-3)
        pm, nst = cvo.perona_malik(ch, 12.0, 0.2, 0.7)
        d[name + "_pm"], d[name + "_pmsteps"] = np.stack(pm), nst
        d[name + "_mask"] = cvo.mask(u)
        d[name + "_params"] = np.array([mu, nu, dt, eps] + lam1 + [0] * (3 - n) + lam2 + [0] * (3 - n))
    d["checker_23x37"] = cvo.levelset_checkerboard(23, 37)
    d["rect_20x30"] = cvo.levelset_rect(20, 30, 4, 5, 11, 7)
    d["circ_40x50"] = cvo.levelset_circ(40, 50, 22, 19, 13)
    d["circ_clip"] = cvo.levelset_circ(30, 30, 3, 27, 9)
    xs = np.concatenate([np.linspace(-40, 40, 2001), rng.standard_normal(200) * 1e3])
    d["hd_x"] = xs
    d["heaviside"] = np.array([cvo.regularized_heaviside(x, 0.7) for x in xs])
    d["delta"] = np.array([cvo.regularized_delta(x, 0.7) for x in xs])
    np.savez_compressed(os.path.join(OUT, "kat_small.npz"), **d)


def config1():
    c = synth.CONFIGS["C1"]
    ch = synth.seastar()
    pm, nst = cvo.perona_malik(ch, **c["pm"])
    u, steps, nrm = cvo.csv_run(pm, cvo.levelset_checkerboard(c["h"], c["w"]), 0.5, 0.0, 1.0, 1.0, [1, 1, 1], [1, 1, 1],
                                1e-3, c["csv"]["max_steps"])
    np.savez_compressed(os.path.join(OUT, "c1.npz"), pm=np.stack(pm), pmsteps=nst, u=u, steps=steps, norm=nrm,
                        mask=np.packbits(cvo.mask(u)))


def config2():
    c = synth.CONFIGS["C2"]
    ch = synth.night_lights()
    pm, nst = cvo.perona_malik(ch, **c["pm"])
    k = c["csv"]
    u, steps, nrm = cvo.csv_run(pm, cvo.levelset_checkerboard(c["h"], c["w"]), 0.5, k["nu"], k["dt"], 1.0, k["lambda1"],
                                [1, 1, 1], k["tol"], k["max_steps"])
    np.savez_compressed(os.path.join(OUT, "c2.npz"), pm=np.stack(pm), pmsteps=nst, u_sub=u[::4, ::4].copy(),
                        u_norm=np.linalg.norm(u), steps=steps, norm=nrm, mask=np.packbits(cvo.mask(u)))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    small_cases()
    config1()
    config2()
    print(sorted(os.listdir(OUT)))
