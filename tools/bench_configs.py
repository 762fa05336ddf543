"""Throughput of the other BASELINE.json configurations (C1, C2, C3, C5) on one GPU -- the record for the results
table in DESIGN.md.  bench.py measures the headline configuration (C4).  Usage: python tools/bench_configs.py [C1 C2 C3 C5]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chan_vese_b200 as cv  # noqa: E402
from chan_vese_b200 import synth  # noqa: E402

PEAK = 6548.2


def timed(fn, reps, stream):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    out = None
    for _ in range(reps):
        out = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


FP32 = os.environ.get("CVB_FP32", "0") == "1"  # the fp32 variant, reported separately


def main():
    which = sys.argv[1:] or ["C1", "C2", "C3", "C5"]
    # C5 across GPUs (torchrun): images are independent, every rank takes count / world of them, no communication
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        which = ["C5"]
    stream = torch.cuda.Stream()
    ctx = cv.Context(local, stream=stream.cuda_stream)
    ctx.set_tile_rows(int(os.environ.get("CVB_TILE_ROWS", "0")))
    res = {}
    for name in which:
        c = synth.CONFIGS[name]
        h, w, n = c["h"], c["w"], c["n"]
        k = dict(c["csv"])
        max_steps = k.pop("max_steps")
        tol = k.pop("tol", 1e-3)
        prm = cv.make_params(nch=n, **k)
        if name == "C5":
            total = int(os.environ.get("C5_COUNT", "4096"))
            count = total // world
            base = synth.batch_images(0, 64, h, w)
            imgs = np.ascontiguousarray(np.tile(base, (count // 64, 1, 1, 1)))
            job = cv.Batch(ctx, count, n, h, w, fp32=FP32)
            job.upload_images(imgs)
            job.save_images()

            def run():
                job.restore_images()
                job.init_checkerboard()
                npm = job.perona_malik(**c["pm"])
                steps, _ = job.csv_run(prm, tol=tol, max_steps=max_steps)
                return npm, steps
            ctx.reset_stats()
            ms, (npm, steps) = timed(run, 2, stream)
            st = ctx.stats()
            pixit = float(h) * w * (npm * count + int(steps.sum()))
            if world > 1:  # whole-job throughput: all ranks' pixel-iterations over the slowest rank's time
                t = torch.tensor([ms, pixit], dtype=torch.float64, device="cuda")
                tmax = t.clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                ms, pixit = float(tmax[0]), float(t[1])
            res[name] = dict(images=count * world, gpus=world, pm_steps=npm, csv_steps_mean=float(steps.mean()), csv_steps_min=int(steps.min()),
                             csv_steps_max=int(steps.max()), ms=ms, pixel_iters_per_s=pixit / (ms * 1e-3))
            job.close()
        else:
            img = {"C1": synth.seastar, "C2": synth.night_lights, "C3": synth.two_phase}[name]()
            sess = cv.Session(ctx, n, h, w, fp32=FP32)
            sess.upload_image(img)
            sess.save_image()
            u0 = cv.levelset_circ(h, w, w // 2, h // 2, h // 4) if c["init"] == "circ" else None

            def run():
                sess.restore_image()
                if u0 is None:
                    sess.init_checkerboard()
                else:
                    sess.upload_levelset(u0)
                npm = sess.perona_malik(**c["pm"]) if c["pm"] else 0
                steps, _ = sess.csv_run(prm, tol=tol, max_steps=max_steps)
                return npm, steps
            ctx.reset_stats()
            ms, (npm, steps) = timed(run, 3, stream)
            st = ctx.stats()
            reps = 4
            csv_ms = st["csv_ms"] / reps / max(steps, 1)
            pm_ms = st["pm_ms"] / reps / max(npm, 1)
            res[name] = dict(pm_steps=npm, csv_steps=steps, ms=ms, pixel_iters_per_s=float(h) * w * (npm + steps) / (ms * 1e-3),
                             csv_us_per_step=csv_ms * 1e3, pm_us_per_step=pm_ms * 1e3,
                             csv_frac_hbm=((8 if FP32 else 16) + n) * h * w / (csv_ms * 1e-3) / 1e9 / PEAK if steps else None,
                             pm_frac_hbm=(8 if FP32 else 16) * n * h * w / (pm_ms * 1e-3) / 1e9 / PEAK if npm else None)
            sess.close()
        if rank == 0:
            print(name, json.dumps(res[name]), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
