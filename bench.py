#!/usr/bin/env python
"""bench.py -- pixel-iterations/s of the fused PM + CSV hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size S]

One "step" = one whole pass of the hot path over the workload: restore the image, device checkerboard init,
Perona-Malik (20 steps) and the Chan-Sandberg-Vese loop (100 steps, tolerance 0) on a 16384 x 16384 RGB image
(BASELINE.json configs[3], the configuration the metric is quoted on; it fits one GPU).  N > 1: the image is cut
into row slabs, one process per GPU (torchrun), NCCL halo exchange + all-gather of the region sums every step.

`value`  : inputs resident in HBM when the timed region starts.
`e2e`    : the same through the reference-facing C ABI with HOST buffers -- every step uploads the uint8 planes
           from pinned memory and reads the segmentation mask back.
`--impl reference`: the reference's CPU path (pass-structured OpenMP port oracle/ref_cpu.cpp; the reference itself
           needs OpenCV 2.4 + Boost and cannot be built here) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pixel-iterations/s of PM+CSV step"
UNIT = "pixel-iterations/s"
PM = dict(K=10.0, L=0.25, T=5.0)   # 20 diffusion steps
CSV_STEPS = 100
BYTES_CSV_RGB = 19.0               # read u 8 + write u 8 + 3 x uint8
BYTES_PM_RGB = 48.0                # (read 8 + write 8) x 3 channels
# dram__bytes_read.sum + dram__bytes_write.sum of csv_step_kernel<3> from the ncu --set full capture
# profiles/r1e_ncu_full_csv_step.csv (8192^2 RGB: 0.784 + 0.508 = 1.292 GB per launch vs 1.275 GB algorithmic; the
# extra 1.3 % are the rows the cp.async ring requests past the end of a segment)
NCU_TRAFFIC_RATIO_CSV = 1.292 / 1.275


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[6]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def cpu_sample(size, steps, warmup):
    """The reference-style CPU path on a bounded sample: a size x size crop of the same scene, PM 2 + CSV 10 steps
    (the workload's 1:5 ratio).  Returns (pixel-iterations/s, description, cores)."""
    import numpy as np  # noqa: F401
    from chan_vese_b200 import synth
    from oracle import coracle as co
    H = W = 16384
    planes = [np.ascontiguousarray(p[:, :size]) for p in synth.hashed_scene_rows(H, W, 0, size, threads=4)]
    u0 = co.levelset_checkerboard(size, size)
    prm = co.params()
    n_pm, n_csv = 2, 10
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        pm, got = co.perona_malik(planes, PM["K"], PM["L"], n_pm * PM["L"], impl="refcpu")
        assert got == n_pm
        _, done, _ = co.csv_run(pm, u0, prm, 0.0, n_csv, impl="refcpu")
        assert done == n_csv
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    value = size * size * (n_pm + n_csv) * len(times) / total
    desc = ("%dx%d crop of the 16384^2 RGB scene, PM %d + CSV %d steps per sample step, oracle/ref_cpu.cpp "
            "(reference pass structure, <=3 OpenMP threads + delta pool)" % (size, size, n_pm, n_csv))
    return value, desc, min(os.cpu_count() or 1, 3), total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, desc, cores, sec = cpu_sample(args.cpu_size, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    s = args.size
    return {"workload": "configs[3]: synthetic %dx%d RGB PM+CSV (PM -K 10 -L 0.25 -T 5 = 20 steps, CSV -N %d -t 0, "
                        "checkerboard init)%s" % (s, s, CSV_STEPS, "" if s == 16384 else " [REDUCED SIZE]"),
            "h": s, "w": s, "channels": 3, "pm_steps": 20, "csv_steps": CSV_STEPS,
            "decomposition": ("row slabs x%d, boundary rows + region sums pushed into peer memory (CUDA IPC over NVLink), "
                              "flag-synchronised; CVB_COMM=nccl switches to NCCL send/recv + all-gather" % world) if world > 1 else "single GPU",
            "l2": "inputs larger than L2 (u ping-pong %.1f GB per GPU)" % (2 * 8.0 * s * s / world / 1e9)}


def run_b200(args):
    import numpy as np
    import torch

    import chan_vese_b200 as cv
    from chan_vese_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.watchdog > 0:
        # a rank that never comes back (a peer lost, a collective stuck) must not hold the box until the caller's own
        # limit: report and leave -- os._exit tears the CUDA context down with the process
        def give_up():
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world,
                                  "error": "watchdog: no result after %d s" % args.watchdog}), flush=True)
            os._exit(3)
        wd = threading.Timer(args.watchdog, give_up)
        wd.daemon = True
        wd.start()
    def trace(msg):  # CVB_BENCH_TRACE=1: one stderr line per phase and rank (where does a multi-rank run stop?)
        if os.environ.get("CVB_BENCH_TRACE"):
            print("[rank %d %.1fs] %s" % (rank, time.time() - t_start, msg), file=sys.stderr, flush=True)

    t_start = time.time()
    h = w = args.size
    stream = torch.cuda.Stream()
    ctx = cv.Context(local, stream=stream.cuda_stream)
    tile_rows = args.tile_rows or cv.auto_tile_rows(h, w, 1, world)
    ctx.set_tile_rows(tile_rows)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(ctx.comm_create_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), world, rank)
        lo, hi = cv.slab_partition(h, tile_rows, world, rank)
        sess = cv.Session(ctx, 3, h, w, rows=(lo, hi))
    else:
        lo, hi = 0, h
        sess = cv.Session(ctx, 3, h, w)
    rows = hi - lo
    # synthetic input in pinned host memory (identical bytes whatever the slab decomposition)
    pinned = [torch.empty((rows, w), dtype=torch.uint8).pin_memory() for _ in range(3)]
    views = [p.numpy() for p in pinned]
    synth.hashed_scene_rows(h, w, lo, hi, out=views, threads=min(8, os.cpu_count() or 1))
    mask_pin = torch.empty((rows, (w + 7) // 8), dtype=torch.uint8).pin_memory()
    mask_view = mask_pin.numpy()
    prm = cv.make_params()
    trace("session created, slab rows [%d, %d), tile rows %d" % (lo, hi, tile_rows))
    sess.upload_image(views)
    sess.save_image()
    trace("image uploaded")

    def step_resident():
        sess.restore_image()
        sess.init_checkerboard()
        n_pm = sess.perona_malik(PM["K"], PM["L"], PM["T"])
        n_csv, _ = sess.csv_run(prm, tol=0.0, max_steps=CSV_STEPS)
        return n_pm, n_csv

    def step_e2e():
        # host buffers in, host buffer out: planes uploaded from pinned memory (later planes behind the diffusion of
        # the earlier ones), the segmentation mask read back bit-packed
        n_pm = sess.upload_image_smooth(views, PM["K"], PM["L"], PM["T"])
        sess.init_checkerboard()
        n_csv, _ = sess.csv_run(prm, tol=0.0, max_steps=CSV_STEPS)
        sess.mask_packed(out=mask_view)
        return n_pm, n_csv

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks):
        for k in range(warmup):
            fn()
            trace("%s: warm-up step %d done" % (fn.__name__, k))
        barrier()
        ctx.reset_stats()
        sampler = ClockSampler(local) if sample_clocks and rank == 0 else None  # one nvidia-smi poller per box
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            n_pm, n_csv = fn()
        e1.record(stream)
        trace("%s: %d timed steps enqueued and returned" % (fn.__name__, steps))
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, n_pm, n_csv, ctx.stats(), clocks

    ms, n_pm, n_csv, st, clocks = timed(step_resident, args.steps, args.warmup, True)
    assert n_pm == 20 and n_csv == CSV_STEPS, (n_pm, n_csv)
    pix_iters = float(h) * w * (n_pm + n_csv) * args.steps
    value = pix_iters / (ms * 1e-3)
    ms_e2e, _, _, st_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 3), False)
    e2e_value = pix_iters / (ms_e2e * 1e-3)
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    csv_launch_ms = st["csv_ms"] / max(st["csv_step_launches"], 1)
    pm_launch_ms = st["pm_ms"] / max(st["pm_step_launches"], 1)
    csv_gbs = BYTES_CSV_RGB * rows * w / (csv_launch_ms * 1e-3) / 1e9
    pm_gbs = BYTES_PM_RGB * rows * w / (pm_launch_ms * 1e-3) / 1e9
    alg_bytes_step = (BYTES_CSV_RGB * n_csv + BYTES_PM_RGB * n_pm) * rows * w
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": st_e2e["h2d_bytes"] // args.steps, "d2h_bytes_per_step": st_e2e["d2h_bytes"] // args.steps},
        "gpu_launches": int(st["kernel_launches"]),
        "roofline": {"bound": "hbm", "kernel": "csv_step_kernel<3,fast>", "achieved": csv_gbs, "peak": peak, "unit": "GB/s",
                     "frac": csv_gbs / peak, "traffic": NCU_TRAFFIC_RATIO_CSV * BYTES_CSV_RGB * rows * w,
                     "traffic_source": "profiles/r1e_ncu_full_csv_step.csv (ncu at 8192^2, scaled by pixels)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_CSV_RGB * rows * w, "launch_ms": csv_launch_ms,
                     "share_of_step": st["csv_ms"] / ms},
        "kernels": {"csv_step": {"launches": int(st["csv_step_launches"]), "ms_per_launch": csv_launch_ms, "GBps": csv_gbs,
                                 "frac": csv_gbs / peak, "pixel_iters_per_s": rows * w / (csv_launch_ms * 1e-3)},
                    "pm_step": {"launches": int(st["pm_step_launches"]), "ms_per_launch": pm_launch_ms, "GBps": pm_gbs,
                                "frac": pm_gbs / peak, "pixel_iters_per_s": rows * w / (pm_launch_ms * 1e-3)},
                    "whole_step_GBps": alg_bytes_step * args.steps / (ms * 1e-3) / 1e9,
                    "whole_step_frac": alg_bytes_step * args.steps / (ms * 1e-3) / 1e9 / peak},
    }
    if world == 1 and not args.no_cpu:
        v, desc, cores, _ = cpu_sample(args.cpu_size, 1, 0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
    print(json.dumps(line), flush=True)
    sess.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=16384, help="image side (default: the BASELINE 16384)")
    ap.add_argument("--cpu-size", type=int, default=1024, help="side of the CPU-baseline crop")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--tile-rows", type=int, default=0, help="rows per tile (0 = the library's automatic choice)")
    ap.add_argument("--watchdog", type=int, default=300, help="give up after this many seconds (0 = never)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
