"""CPU, world_size 2 over gloo: the row-slab decomposition the multi-GPU path uses (cvb_slab_partition, 2-row halos
each way, all-gather of per-slab region sums) reproduces the whole-image result when every rank computes its slab
with the oracle.  The CUDA kernels are not involved (no GPU here); tools/multigpu_check.py is the GPU counterpart."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import chan_vese_b200 as cv
from chan_vese_b200 import synth
from oracle import coracle as co

H, W, ROWS = 96, 70, 8
HALO = 2
PM_STEPS, CSV_STEPS = 3, 4
K, L = 15.0, 0.25


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _exchange(x, lo, hi, rank, world):
    """x: (hi-lo+2*HALO, W) with the slab in rows [HALO, HALO+hi-lo): fill halo rows from the neighbours."""
    n = hi - lo
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(x[HALO:2 * HALO].copy()), rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(x[n:n + HALO].copy()), rank + 1))
    if rank > 0:
        t = torch.empty((HALO, x.shape[1]), dtype=torch.from_numpy(x).dtype)
        dist.recv(t, rank - 1)
        x[0:HALO] = t.numpy()
    if rank < world - 1:
        t = torch.empty((HALO, x.shape[1]), dtype=torch.from_numpy(x).dtype)
        dist.recv(t, rank + 1)
        x[n + HALO:n + 2 * HALO] = t.numpy()
    for r in reqs:
        r.wait()


def _ext(x, lo, hi, rank, world):
    """rows of the halo-extended slab that exist in the image (drop the unused halo at the image border)."""
    a = 0 if rank > 0 else HALO
    b = x.shape[0] if rank < world - 1 else x.shape[0] - HALO
    return a, b


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = cv.slab_partition(H, ROWS, world, rank)
    n = hi - lo
    img = synth.hashed_scene_rows(H, W, lo, hi, cell=32)
    # ---- PM: fp64 state per channel, 2 halo rows each way every step, quantise at the end
    planes = []
    for ch in img:
        x = np.zeros((n + 2 * HALO, W))
        x[HALO:HALO + n] = ch
        for _ in range(PM_STEPS):
            _exchange(x, lo, hi, rank, world)
            a, b = _ext(x, lo, hi, rank, world)
            y = co.pm_evolve(x[a:b], K, L, 1)
            x[HALO:HALO + n] = y[HALO - a:HALO - a + n]
        planes.append(np.clip(np.rint(x[HALO:HALO + n]), 0, 255).astype(np.uint8))
    # ---- CSV: sums all-gathered every step, u halos exchanged every step
    u = np.zeros((n + 2 * HALO, W))
    u[HALO:HALO + n] = co.levelset_checkerboard(H, W)[lo:hi]
    imgx = []
    for p in planes:
        x = np.zeros((n + 2 * HALO, W), dtype=np.uint8)
        x[HALO:HALO + n] = p
        _exchange(x, lo, hi, rank, world)
        imgx.append(x)
    prm = co.params(lambda1=[1.0, 0.5, 2.0])
    for _ in range(CSV_STEPS):
        hv = np.array([co.oracle().cvo_heaviside(v, 1.0) for v in u[HALO:HALO + n].ravel()]).reshape(n, W)
        sums = [hv.sum(), (1 - hv).sum()] + [(p * hv).sum() for p in planes] + [(p * (1 - hv)).sum() for p in planes]
        mine = torch.tensor(sums, dtype=torch.float64)
        allsums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allsums, mine)
        tot = torch.stack(allsums).sum(0).numpy()  # fixed rank order on every rank
        c1, c2 = tot[2:5] / tot[0], tot[5:8] / tot[1]
        _exchange(u, lo, hi, rank, world)
        a, b = _ext(u, lo, hi, rank, world)
        un, _, _, _ = co.csv_step([x[a:b] for x in imgx], u[a:b], prm, c1, c2)
        u[HALO:HALO + n] = un[HALO - a:HALO - a + n]
    np.save(os.path.join(out, "u_%d.npy" % rank), u[HALO:HALO + n])
    np.save(os.path.join(out, "pm_%d.npy" % rank), np.stack(planes))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slabs_equal_whole_image(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    u = np.concatenate([np.load(tmp_path / ("u_%d.npy" % r)) for r in range(world)])
    pm = np.concatenate([np.load(tmp_path / ("pm_%d.npy" % r)) for r in range(world)], axis=1)
    whole = synth.hashed_scene_rows(H, W, 0, H, cell=32)
    ref_pm, nst = co.perona_malik(whole, K, L, PM_STEPS * L)
    assert nst == PM_STEPS
    assert np.array_equal(pm, np.stack(ref_pm))  # stencil radius 2 per step == 2 halo rows: exact
    ref_u, steps, _ = co.csv_run(ref_pm, co.levelset_checkerboard(H, W), co.params(lambda1=[1.0, 0.5, 2.0]), 0.0, CSV_STEPS)
    assert steps == CSV_STEPS
    assert np.linalg.norm(u - ref_u) / np.linalg.norm(ref_u) < 1e-12  # only the summation order of c1/c2 differs


def test_slab_spans_match_group_ownership():
    # a slab is a union of whole reduction groups: rank r of G owns groups [32r/G, 32(r+1)/G)
    for h, rows in [(16384, 128), (4096, 32), (1000, 8)]:
        nseg = -(-h // rows)
        for world in (2, 4, 8):
            for r in range(world):
                lo, hi = cv.slab_partition(h, rows, world, r)
                g0, g1 = r * 32 // world, (r + 1) * 32 // world
                assert lo == min(-(-g0 * nseg // 32) * rows, h) and hi == min(-(-g1 * nseg // 32) * rows, h)


# ---- bench.py's result digest across ranks (gloo, world_size 2): per-rank CRCs gathered with all_gather_object and
# combined in rank order equal the CRC of the whole image, for the automatic (rank-count independent) tile length ----------
def _digest_worker(rank, world, port, out):
    import sys
    import zlib
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    h, w = 9000, 4000
    tile = cv.auto_tile_rows(h, w, 1, world)
    assert tile == cv.auto_tile_rows(h, w, 1, 1)  # the same tiling on every rank count
    lo, hi = cv.slab_partition(h, tile, world, rank)
    planes = synth.hashed_scene_rows(h, w, lo, hi, cell=512)
    u = np.cos(np.arange(lo, hi, dtype=np.float64))[:, None] * np.arange(w, dtype=np.float64)[None, :]
    parts = [bench.crc_of(p) for p in planes] + [bench.crc_of(u)]
    gathered = [None] * world
    dist.all_gather_object(gathered, parts)
    if rank == 0:
        whole = synth.hashed_scene_rows(h, w, 0, h, cell=512)
        uw = np.cos(np.arange(h, dtype=np.float64))[:, None] * np.arange(w, dtype=np.float64)[None, :]
        want = [zlib.crc32(p.tobytes()) for p in whole] + [zlib.crc32(uw.tobytes())]
        out.put(bench.combine_ranks(gathered) == want)
    dist.barrier()
    dist.destroy_process_group()


def test_bench_digest_over_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_digest_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert ok and all(p.exitcode == 0 for p in procs)
