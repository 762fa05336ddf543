"""Deterministic synthetic inputs for the five BASELINE.json configurations.

The reference's own input images are not in its tree (README.md:49,69 link to Wikimedia), so every parity test
and benchmark runs on generated images: planar uint8 B,G,R planes, identical bytes for the oracle and the GPU.
C1-C3 use numpy's PCG64; C4/C5 use a counter-based hash of the global pixel index so that every rank can
generate its own row slab / its own images without communication.
"""
import math

import numpy as np

# parameters of the two README command lines (README.md:53, 59-62) and of the synthetic studies
CONFIGS = {
    "C1": dict(h=250, w=370, n=3, pm=dict(K=30.0, L=0.25, T=100.0), csv=dict(max_steps=70), init="checkerboard"),
    "C2": dict(h=430, w=640, n=3, pm=dict(K=1000.0, L=0.1, T=1.5),
               csv=dict(max_steps=132, dt=0.001, tol=1e-6, nu=-293.0, lambda1=[1.0, 1.0, 0.1]), init="checkerboard"),
    "C3": dict(h=4096, w=4096, n=1, pm=None, csv=dict(max_steps=2000, tol=0.0), init="circ"),
    "C4": dict(h=16384, w=16384, n=3, pm=dict(K=10.0, L=0.25, T=5.0), csv=dict(max_steps=100, tol=0.0),
               init="checkerboard"),
    "C5": dict(h=512, w=512, n=3, count=4096, pm=dict(K=30.0, L=0.25, T=10.0), csv=dict(max_steps=50),
               init="checkerboard"),
}


def _clip_u8(x):
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def seastar(h=250, w=370, seed=1, arms=11, inside=(40, 110, 220), outside=(150, 120, 60), sigma=25.0):
    """C1: an `arms`-armed star r < 0.22h + 0.12h cos(arms*theta), B,G,R fore/background colours, N(0, sigma) noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    dy, dx = yy - (h - 1) / 2.0, xx - (w - 1) / 2.0
    r = np.hypot(dx, dy)
    th = np.arctan2(dy, dx)
    star = r < 0.22 * h + 0.12 * h * np.cos(arms * th)
    planes = []
    for k in range(3):
        base = np.where(star, float(inside[k]), float(outside[k]))
        planes.append(_clip_u8(base + rng.normal(0.0, sigma, size=(h, w))))
    return planes


def night_lights(h=430, w=640, seed=2, blobs=400):
    """C2: dark background N(8,4) with `blobs` Gaussian lights clustered in the middle half, colour weights B .5 G .85 R 1."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    lum = np.zeros((h, w))
    cy = rng.uniform(0.25 * h, 0.75 * h, blobs)
    cx = rng.uniform(0.25 * w, 0.75 * w, blobs)
    sg = rng.uniform(1.0, 6.0, blobs)
    am = rng.uniform(80.0, 255.0, blobs)
    for b in range(blobs):
        y0, y1 = max(0, int(cy[b] - 5 * sg[b])), min(h, int(cy[b] + 5 * sg[b]) + 1)
        x0, x1 = max(0, int(cx[b] - 5 * sg[b])), min(w, int(cx[b] + 5 * sg[b]) + 1)
        d2 = (yy[y0:y1, x0:x1] - cy[b]) ** 2 + (xx[y0:y1, x0:x1] - cx[b]) ** 2
        lum[y0:y1, x0:x1] += am[b] * np.exp(-d2 / (2 * sg[b] ** 2))
    planes = []
    for wk in (0.5, 0.85, 1.0):
        planes.append(_clip_u8(wk * lum + rng.normal(8.0, 4.0, size=(h, w))))
    return planes


def two_phase(h=4096, w=4096, seed=3, discs=24, lo=70.0, hi=180.0, sigma=20.0):
    """C3: grayscale union of discs at `hi` on `lo` plus N(0, sigma) noise."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w), lo, dtype=np.float32)
    cy = rng.uniform(0, h, discs)
    cx = rng.uniform(0, w, discs)
    rr = rng.uniform(0.03, 0.12, discs) * min(h, w)
    for b in range(discs):
        y0, y1 = max(0, int(cy[b] - rr[b])), min(h, int(cy[b] + rr[b]) + 1)
        x0, x1 = max(0, int(cx[b] - rr[b])), min(w, int(cx[b] + rr[b]) + 1)
        yy, xx = np.ogrid[y0:y1, x0:x1]
        img[y0:y1, x0:x1][(yy - cy[b]) ** 2 + (xx - cx[b]) ** 2 < rr[b] ** 2] = hi
    img += rng.normal(0.0, sigma, size=(h, w)).astype(np.float32)
    return [_clip_u8(img)]


def _hash32(x):
    """lowbias32-style integer hash on uint32 arrays (wraps modulo 2^32)."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def hashed_scene_rows(h, w, row_lo, row_hi, seed=4, n=3, cell=1024, chunk=256, out=None, threads=1):
    """C4: rows [row_lo, row_hi) of an h x w scene of one disc per `cell`-pixel cell (centre, radius and colour from a
    hash of the cell index) plus hash noise; depends only on the GLOBAL pixel index, so slabs agree with the whole.
    `out`: optional list of n preallocated (rows, w) uint8 arrays (e.g. pinned memory)."""
    if out is None:
        out = [np.empty((row_hi - row_lo, w), dtype=np.uint8) for _ in range(n)]
    jj = np.arange(w, dtype=np.int64)
    cj = jj // cell
    fg = (210.0, 160.0, 90.0)
    bg = (60.0, 95.0, 140.0)

    def work(r0):
        r1 = min(r0 + chunk, row_hi)
        ii = np.arange(r0, r1, dtype=np.int64)
        ci = ii // cell
        cid = (ci[:, None] * 65537 + cj[None, :] + seed * 7919).astype(np.uint32)
        hc = _hash32(cid)
        cy = ci[:, None] * cell + cell // 4 + (hc & np.uint32(0xFF)).astype(np.int64) * (cell // 2) // 256
        cx = cj[None, :] * cell + cell // 4 + ((hc >> np.uint32(8)) & np.uint32(0xFF)).astype(np.int64) * (cell // 2) // 256
        rad = cell // 8 + ((hc >> np.uint32(16)) & np.uint32(0xFF)).astype(np.int64) * (cell // 6) // 256
        inside = (ii[:, None] - cy) ** 2 + (jj[None, :] - cx) ** 2 < rad ** 2
        pix = (ii[:, None] * w + jj[None, :])
        for k in range(n):
            hb = _hash32((pix * n + k + seed * 104729).astype(np.uint32))
            noise = ((hb & np.uint32(0xFF)).astype(np.float32) + ((hb >> np.uint32(8)) & np.uint32(0xFF)).astype(np.float32) +
                     ((hb >> np.uint32(16)) & np.uint32(0xFF)).astype(np.float32) + (hb >> np.uint32(24)).astype(np.float32) -
                     510.0) * np.float32(0.125)
            base = np.where(inside, np.float32(fg[k % 3]), np.float32(bg[k % 3]))
            out[k][r0 - row_lo:r1 - row_lo] = _clip_u8(base + noise)

    starts = list(range(row_lo, row_hi, chunk))
    if threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, starts))
    else:
        for r0 in starts:
            work(r0)
    return out


def batch_image(index, h=512, w=512, seed=5):
    """C5: image `index` of the batch -- a C1-style star with hash-chosen arm count and colours."""
    s = (seed << 32) + index
    rng = np.random.default_rng(s)
    arms = int(rng.integers(5, 13))
    inside = tuple(int(v) for v in rng.integers(20, 236, 3))
    outside = tuple(int((v + 128) % 256) for v in inside)
    return seastar(h, w, seed=s + 1, arms=arms, inside=inside, outside=outside, sigma=20.0)


def batch_images(first, count, h=512, w=512, seed=5):
    out = np.empty((count, 3, h, w), dtype=np.uint8)
    for m in range(count):
        planes = batch_image(first + m, h, w, seed)
        for k in range(3):
            out[m, k] = planes[k]
    return out


def pm_steps_expected(L, T):
    """Python replay of `for (double t = 0; t < T; t += L)` (src/main.cpp:498); used by tests as an independent count."""
    n, t = 0, 0.0
    while t < T:
        t += L
        n += 1
    return n


def checkerboard_sign_vectors(h, w):
    si = [math.sin(math.pi * i / 5) for i in range(h)]
    sj = [math.sin(math.pi * j / 5) for j in range(w)]
    return si, sj
