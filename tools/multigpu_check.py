"""Row-slab run over N GPUs (torchrun, one process per GPU) against the single-GPU run of the same image:
PM planes, level set, step count and norm must be BIT-IDENTICAL (fixed reduction groups).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py [--size 2048]

The geometry of bench.py (the open 8-GPU item of DESIGN.md section 5) is reproduced with
    ... tools/multigpu_check.py --size 16384 --square --tiles world --pm-T 5 --csv-steps 100 --repeat 3 --trace
(tiles chosen for the rank count as bench.py does, 20 PM + 100 CSV steps, the resident sequence with save / restore
repeated, the end-to-end sequence with the bit-packed mask, a trace line per phase and rank on stderr).
Programmatic dependent launch and the overlapped upload are on at every rank count (CVB_PDL=0 / CVB_OVERLAP_UPLOAD=0 turn
them off).  Soak: `--repeat 50` repeats the resident sequence 50 times before the comparison (round 2: PASS on 8 GPUs).
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chan_vese_b200 as cv  # noqa: E402
from chan_vese_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--csv-steps", type=int, default=12)
    ap.add_argument("--square", action="store_true", help="w = size (default: size - 200, a ragged last strip)")
    ap.add_argument("--tiles", choices=["one", "world"], default="one", help="tile rows chosen for 1 GPU or for the rank count")
    ap.add_argument("--pm-T", type=float, default=1.5, help="Perona-Malik duration (L = 0.25: 4 steps per unit)")
    ap.add_argument("--repeat", type=int, default=1, help="repeat the resident sequence (save / restore in between)")
    ap.add_argument("--trace", action="store_true", help="one stderr line per phase and rank")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h, w = args.size, args.size if args.square else args.size - 200

    def trace(msg):
        if args.trace:
            print("[rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    ctx = cv.Context(local)
    rows = cv.auto_tile_rows(h, w, 1, world if args.tiles == "world" else 1)
    ctx.set_tile_rows(rows)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(ctx.comm_create_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.comm_init(idt.cpu().numpy().tobytes(), world, rank)
    lo, hi = cv.slab_partition(h, rows, world, rank)
    img = synth.hashed_scene_rows(h, w, lo, hi, cell=256, threads=4)
    prm = cv.make_params(lambda1=[1.0, 0.5, 2.0])
    trace("slab rows [%d, %d), tile rows %d" % (lo, hi, rows))
    with cv.Session(ctx, 3, h, w, rows=(lo, hi)) as s:
        trace("session created")
        s.upload_image(img)
        s.save_image()
        for it in range(args.repeat):  # bench.py's resident sequence
            s.restore_image()
            s.init_checkerboard()
            n_pm = s.perona_malik(20.0, 0.25, args.pm_T)
            trace("iteration %d: perona_malik done (%d steps)" % (it, n_pm))
            if it + 1 < args.repeat:
                s.csv_run(prm, tol=0.0, max_steps=args.csv_steps)
                trace("iteration %d: csv_run done" % it)
        pm = s.download_image()
        # the overlapped upload + PM path must give the same planes (bench.py's end-to-end sequence)
        n_pm2 = s.upload_image_smooth(img, 20.0, 0.25, args.pm_T)
        trace("upload_image_smooth done")
        pm2 = s.download_image()
        assert n_pm2 == n_pm and all(np.array_equal(a, b) for a, b in zip(pm, pm2)), "upload_image_smooth differs on rank %d" % rank
        steps, norm = s.csv_run(prm, tol=0.0, max_steps=args.csv_steps)
        trace("csv_run done (%d steps)" % steps)
        u = s.download_levelset()
        packed = s.mask_packed()
        assert np.array_equal(np.unpackbits(packed, axis=1)[:, :w].astype(bool), u.astype(np.float32) > 0), "mask_packed differs"
        trace("mask_packed done")
        # and an early-stopping run: every rank must stop at the same step
        s.init_checkerboard()
        steps2, norm2 = s.csv_run(prm, tol=0.15, max_steps=200)
        u2 = s.download_levelset()
    np.save("/tmp/slab_u_%d.npy" % rank, u)
    np.save("/tmp/slab_u2_%d.npy" % rank, u2)
    np.save("/tmp/slab_pm_%d.npy" % rank, np.stack(pm))
    meta = torch.tensor([steps, steps2, n_pm], dtype=torch.int64, device="cuda")
    allmeta = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(allmeta, meta)
    dist.barrier()
    if rank == 0:
        assert all(torch.equal(m, allmeta[0]) for m in allmeta), allmeta
        full_u = np.concatenate([np.load("/tmp/slab_u_%d.npy" % r) for r in range(world)])
        full_u2 = np.concatenate([np.load("/tmp/slab_u2_%d.npy" % r) for r in range(world)])
        full_pm = np.concatenate([np.load("/tmp/slab_pm_%d.npy" % r) for r in range(world)], axis=1)
        ctx1 = cv.Context(local)
        ctx1.set_tile_rows(rows)
        whole = synth.hashed_scene_rows(h, w, 0, h, cell=256, threads=4)
        with cv.Session(ctx1, 3, h, w) as s:
            s.upload_image(whole)
            s.init_checkerboard()
            s.perona_malik(20.0, 0.25, args.pm_T)
            pm1 = np.stack(s.download_image())
            st1, nrm1 = s.csv_run(prm, tol=0.0, max_steps=args.csv_steps)
            u1 = s.download_levelset()
            s.init_checkerboard()
            st2, nrm2 = s.csv_run(prm, tol=0.15, max_steps=200)
            u21 = s.download_levelset()
        ok = (np.array_equal(full_pm, pm1) and np.array_equal(full_u, u1) and steps == st1 and norm == nrm1 and
              np.array_equal(full_u2, u21) and steps2 == st2 and norm2 == nrm2)
        print("MULTIGPU_CHECK world=%d size=%dx%d pm_equal=%s u_equal=%s steps=%d/%d norm_equal=%s early_stop steps=%d/%d u_equal=%s => %s" % (
            world, h, w, np.array_equal(full_pm, pm1), np.array_equal(full_u, u1), steps, st1, norm == nrm1, steps2, st2,
            np.array_equal(full_u2, u21), "PASS" if ok else "FAIL"), flush=True)
        if not ok:
            d = np.abs(full_u - u1)
            print("max abs diff", d.max(), "rows with diff", np.unique(np.nonzero(d)[0])[:20])
    dist.barrier()
    ctx.comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
