"""ORACLE (test infrastructure, not product): call-for-call Python/cv2 transcription of the
reference hot path.

The reference (ktht/chan_vese, C++14) delegates its per-pixel array arithmetic to OpenCV 2.4.8
(README.md:12), which is not vendored under /root/reference and whose C++ headers are absent here.
The one OpenCV that *is* available is the Python binding cv2 4.13; this file calls the same OpenCV
entry points at the same call sites as the reference so that the plain-C restatement in
`oracle/cv_oracle.c` can be validated against real OpenCV arithmetic, and so that golden fixtures
can be produced (`oracle/make_golden.py`).

PARITY UNPINNED by the reference itself: the reference has no tests, golden vectors or fixtures
(SURVEY.md section 4).  What pins this oracle is OpenCV 4.13's behaviour on the reference's call sites.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Race-free reading of the reference's two data races: intensity_avg starts at zero,
channels are accumulated serially k = 0..N-1.
"""
import math

import cv2
import numpy as np

# src/main.cpp:120-125 -- finite-difference kernels (cv::filter2D = correlation, centre anchor)
FWD_X = np.array([[0.0, -1.0, 1.0]])
FWD_Y = FWD_X.T.copy()
BWD_X = np.array([[-1.0, 1.0, 0.0]])
BWD_Y = BWD_X.T.copy()
CTR_X = np.array([[-0.5, 0.0, 0.5]])
CTR_Y = CTR_X.T.copy()


def regularized_heaviside(x, eps=1.0):
    """src/main.cpp:188-194."""
    return (1 + 2 / math.pi * math.atan(x / eps)) / 2


def regularized_delta(x, eps=1.0):
    """src/main.cpp:204-210 (std::pow(x, 2) == x*x)."""
    return eps / (math.pi * (eps * eps + x * x))


def levelset_checkerboard(h, w):
    """src/main.cpp:221-233: sign(sin(pi*i/5) * sin(pi*j/5)); glibc sin via math.sin."""
    u = np.empty((h, w), dtype=np.float64)
    si = [math.sin(math.pi * i / 5) for i in range(h)]
    sj = [math.sin(math.pi * j / 5) for j in range(w)]
    for i in range(h):
        for j in range(w):
            p = si[i] * sj[j]
            u[i, j] = (p > 0) - (p < 0)
    return u


def levelset_rect(h, w, x, y, rw, rh):
    """src/InteractiveDataRect.cpp:20-27: zeros, u(roi) = 1."""
    u = np.zeros((h, w), dtype=np.float64)
    u[y:y + rh, x:x + rw] = 1.0
    return u


def levelset_circ(h, w, cx, cy, radius):
    """src/InteractiveDataCirc.cpp:18-25: cv::circle(u, P1, radius, 1) -- thickness 1 ring."""
    u = np.zeros((h, w), dtype=np.float64)
    cv2.circle(u, (int(cx), int(cy)), int(radius), 1)
    return u


_ATAN = np.frompyfunc(math.atan, 1, 1)


def heaviside_array(u, eps):
    """regularized_heaviside over an array, each element through glibc atan (math.atan)."""
    return (1 + 2 / math.pi * _ATAN(u / eps).astype(np.float64)) / 2


def region_variance(img, u, inside, eps):
    """src/main.cpp:255-281: serial row-major fp64 sums (np.cumsum accumulates strictly in order)."""
    hv = heaviside_array(u.ravel(), eps)
    if not inside:
        hv = 1 - hv
    nom = np.cumsum(img.ravel().astype(np.float64) * hv)[-1]
    denom = np.cumsum(hv)[-1]
    with np.errstate(all="ignore"):
        return float(np.float64(nom) / np.float64(denom))


def variance_penalty(channel, c, lam):
    """src/main.cpp:299-312: convertTo, -= c, pow 2, *= lambda."""
    t = channel.astype(np.float64)
    t = t - c
    t = cv2.pow(t, 2)
    t = t * lam
    return t


def curvature(u):
    """src/main.cpp:342-375."""
    eta2 = 1e-8 ** 2  # std::pow(1E-8, 2)
    upx = cv2.filter2D(u, cv2.CV_64F, FWD_X, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REPLICATE)
    upy = cv2.filter2D(u, cv2.CV_64F, FWD_Y, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REPLICATE)
    ucx = cv2.filter2D(u, cv2.CV_64F, CTR_X, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REPLICATE)
    ucy = cv2.filter2D(u, cv2.CV_64F, CTR_Y, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REPLICATE)
    nx = upx / np.sqrt(upx * upx + ucx * ucx + eta2)
    ny = upy / np.sqrt(upy * upy + ucy * ucy + eta2)
    kx = cv2.filter2D(nx, cv2.CV_64F, BWD_X, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REPLICATE)
    ky = cv2.filter2D(ny, cv2.CV_64F, BWD_Y, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REPLICATE)
    return kx + ky


def pm_num_steps(L, T):
    """src/main.cpp:498: for (double t = 0; t < T; t += L)."""
    n = 0
    t = 0.0
    while t < T:
        n += 1
        t += L
    return n


def perona_malik(channels, K, L, T):
    """src/main.cpp:478-560, per channel."""
    out = []
    nsteps = pm_num_steps(L, T)
    for ch in channels:
        h, w = ch.shape
        I = ch.astype(np.float64)
        for _ in range(nsteps):
            dx = cv2.Sobel(I, cv2.CV_64F, 1, 0, ksize=3)
            dy = cv2.Sobel(I, cv2.CV_64F, 0, 1, ksize=3)
            g = 1.0 / (1.0 + (dx * dx + dy * dy) / (K * K))
            g[0, :] = 1.0
            g[h - 1, :] = 1.0
            g[:, 0] = 1.0
            g[:, w - 1] = 1.0
            Ip = np.pad(I, 1, mode="edge")
            gp = np.pad(g, 1, mode="edge")
            Is, Ie, In, Iw = Ip[2:, 1:-1], Ip[1:-1, 2:], Ip[:-2, 1:-1], Ip[1:-1, :-2]
            cs, ce, cn, cw = gp[2:, 1:-1], gp[1:-1, 2:], gp[:-2, 1:-1], gp[1:-1, :-2]
            I = I + L * ((cs + g) * (Is - I) + (ce + g) * (Ie - I) + (cn + g) * (In - I) + (cw + g) * (Iw - I)) / 4
        out.append(convert_to_u8(I))
    return out, nsteps


def convert_to_u8(I):
    """Mat::convertTo(CV_8UC1) = saturate_cast<uchar>(cvRound(x)): round-half-even, clamp to 0..255.
    cv2.convertScaleAbs would take |x| first, so the plain conversion goes through cv2.add with a
    zero addend and dtype=CV_8U, which runs OpenCV's own saturate_cast<uchar>(double)."""
    return cv2.add(I, np.zeros_like(I), dtype=cv2.CV_8U)


def stop_condition(channels, tol):
    """src/main.cpp:949-960 (zero-initialised, serial): tol * || (sum_k I_k) * (1/N) ||_2."""
    n = len(channels)
    acc = np.zeros(channels[0].shape, dtype=np.float64)
    for ch in channels:
        acc = acc + ch.astype(np.float64)
    acc = acc * (1.0 / n)  # cv::Mat::operator/=(double) multiplies by the reciprocal
    return tol * cv2.norm(acc, cv2.NORM_L2)


def csv_step(channels, u, mu, nu, dt, eps, lambda1, lambda2):
    """One iteration of src/main.cpp:963-994; returns (u_new, norm, c1, c2)."""
    n = len(channels)
    u_diff = np.zeros(u.shape, dtype=np.float64)
    c1s, c2s = [], []
    for k in range(n):
        c1 = region_variance(channels[k], u, True, eps)
        c2 = region_variance(channels[k], u, False, eps)
        c1s.append(c1)
        c2s.append(c2)
        vi = variance_penalty(channels[k], c1, lambda1[k])
        vo = variance_penalty(channels[k], c2, lambda2[k])
        u_diff = u_diff + (-vi + vo)
    kappa = curvature(u)
    # MatExpr folding of :985 -> one addWeighted: kappa*(mu*dt) + u_diff*((1/N)*dt) + (-nu*dt)
    u_diff = cv2.addWeighted(kappa, mu * dt, u_diff, (1.0 / n) * dt, (-nu) * dt)
    d = eps / (math.pi * (eps * eps + u * u))
    u_diff = cv2.multiply(u_diff, d)
    norm = cv2.norm(u_diff, cv2.NORM_L2)
    return u + u_diff, norm, c1s, c2s


def csv_run(channels, u, mu, nu, dt, eps, lambda1, lambda2, tol, max_steps):
    """src/main.cpp:949-1001; returns (u, steps_done, last_norm)."""
    stop = stop_condition(channels, tol)
    steps = 0
    norm = float("nan")
    for t in range(1, max_steps + 1):
        u, norm, _, _ = csv_step(channels, u, mu, nu, dt, eps, lambda1, lambda2)
        steps = t
        if norm <= stop:
            break
    return u, steps, norm


def mask(u, invert=False):
    """src/main.cpp:395-400: float32(u) > 0 -> 1, optional 1 - mask."""
    _, m = cv2.threshold(u.astype(np.float32), 0, 1, cv2.THRESH_BINARY)
    m = m.astype(np.uint8)
    return (1 - m) if invert else m
