#!/usr/bin/env python
"""bench.py -- pixel-iterations/s of the fused PM + CSV hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload slabs|batch] [--size S]

`--workload slabs` (default, BASELINE.json configs[3], the configuration the metric is quoted on; it fits one GPU):
one "step" = one whole pass of the hot path over a 16384 x 16384 RGB image -- restore the image (device copy), device
checkerboard init, Perona-Malik (20 steps), the Chan-Sandberg-Vese loop (100 steps, tolerance 0).  N > 1: the image is
cut into row slabs, one process per GPU (torchrun); boundary rows and the region sums travel between the GPUs inside
the step kernels (stores into peer memory mapped with CUDA IPC over NVLink, flag-synchronised; `CVB_COMM=nccl` switches
to NCCL send/recv + all-gather between the launches).  Strong scaling: the job is the same for every N, and so is the
result -- `result_digest` (step count, last norm, CRC32 of the PM planes, of the level set and of the packed mask,
combined over the ranks in row order) is identical for N = 1, 2, 4, 8.

`--workload batch` (configs[4]): 4096 independent 512 x 512 RGB images, PM 40 + CSV <= 50 steps with per-image early
stop, images split evenly over the ranks, no communication; strong scaling over the fixed batch.

`value`  : inputs resident in HBM when the timed region starts.
`e2e`    : the same through the C ABI with HOST buffers, as a stream of images through one session -- every step uploads
           its uint8 planes from pinned memory (cvb_session_prefetch_image: the copy of step k+1's planes runs on a second
           stream behind the solver of step k) and reads the bit-packed segmentation mask back.
`extra`  : (N = 1, slabs) the other BASELINE configurations timed in the same run -- C1, C2 whole-job ms, C3 (4096^2
           gray, 2000 steps) with its own roofline fraction, the fp32 variant of C3, 512 of C5's images as one batch --
           and `e2e_oneshot`: one call of
           cvb_segment (the seam INTEGRATION.md binds) with pageable host buffers, fp64 u in and out.
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
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pixel-iterations/s of PM+CSV step"
UNIT = "pixel-iterations/s"
PM = dict(K=10.0, L=0.25, T=5.0)   # 20 diffusion steps
CSV_STEPS = 100
BYTES_CSV_RGB = 19.0               # read u 8 + write u 8 + 3 x uint8
BYTES_PM_RGB = 48.0                # (read 8 + write 8) x 3 channels


def ncu_traffic():
    """DRAM bytes per pixel of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum of one launch, and the pixels that launch processed)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)["csv_step_rgb"]
        return float(t["dram_bytes_per_launch"]) / float(t["pixels_per_launch"]), t["source"]
    except Exception:
        return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CRC32 of a byte stream that is spread over the ranks (zlib's crc32_combine, GF(2) matrix squaring) -------------
def _gf2_times(mat, vec):
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, mat[n]) for n in range(32)]


def crc32_combine(crc1, crc2, len2):
    """CRC32 of A + B from crc32(A), crc32(B) and len(B)."""
    if len2 <= 0:
        return crc1
    odd = [0xEDB88320] + [1 << n for n in range(31)]
    even = _gf2_square(odd)
    odd = _gf2_square(even)
    while True:
        even = _gf2_square(odd)
        if len2 & 1:
            crc1 = _gf2_times(even, crc1)
        len2 >>= 1
        if not len2:
            break
        odd = _gf2_square(even)
        if len2 & 1:
            crc1 = _gf2_times(odd, crc1)
        len2 >>= 1
        if not len2:
            break
    return crc1 ^ crc2


def crc_of(arr):
    import numpy as np
    a = np.ascontiguousarray(arr)
    return zlib.crc32(memoryview(a).cast("B")), a.nbytes


def combine_ranks(parts):
    """parts: per rank (in row order) a list of (crc, nbytes), one entry per stream -> list of combined CRCs."""
    out = []
    for k in range(len(parts[0])):
        crc, _ = parts[0][k]
        for r in range(1, len(parts)):
            crc = crc32_combine(crc, parts[r][k][0], parts[r][k][1])
        out.append(crc)
    return out


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
    (the workload's 1:5 ratio).  Returns (pixel-iterations/s, description, cores, seconds per sample step)."""
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
    # `config` is the b200 arm's config of the same command line, key for key (the measurement contract: the reference
    # arm runs "on your arm's config"); WHAT of that workload a step of this arm times is said next to it, in `sample`
    cfg = workload_config(args, max(args.gpus, 1))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "sample": {"of": cfg["workload"], "h": args.cpu_size, "w": args.cpu_size, "pm_steps": 2, "csv_steps": 10,
                       "what": "every step of this arm is a %dx%d crop of the workload's scene, PM 2 + CSV 10 steps (the "
                               "workload's 1:5 ratio), on the host cores only; the value is per pixel-iteration, so it compares "
                               "with the b200 arm's (a smaller working set favours the CPU: the ratio is conservative)"
                               % (args.cpu_size, args.cpu_size)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    if args.workload == "batch":
        return {"workload": "configs[4]: batch of %d independent 512x512 RGB images, PM (-K 30 -L 0.25 -T 10 = 40 steps) + CSV "
                            "(-N 50, default tolerance: per-image early stop), checkerboard init" % args.batch,
                "images": args.batch, "h": 512, "w": 512, "channels": 3, "pm_steps": 40, "csv_max_steps": 50,
                "decomposition": "images split evenly over %d rank(s), no communication" % world,
                "l2": "inputs larger than L2 (%.1f GB of level-set planes per GPU)" % (2 * 8.0 * 512 * 512 * args.batch / world / 1e9)}
    s = args.size
    return {"workload": "configs[3]: synthetic %dx%d RGB PM+CSV (PM -K 10 -L 0.25 -T 5 = 20 steps, CSV -N %d -t 0, "
                        "checkerboard init)%s" % (s, s, CSV_STEPS, "" if s == 16384 else " [REDUCED SIZE]"),
            "h": s, "w": s, "channels": 3, "pm_steps": 20, "csv_steps": CSV_STEPS,
            "decomposition": ("row slabs x%d, boundary rows + region sums pushed into peer memory (CUDA IPC over NVLink), "
                              "flag-synchronised; CVB_COMM=nccl switches to NCCL send/recv + all-gather" % world) if world > 1 else "single GPU",
            "l2": "inputs larger than L2 (u ping-pong %.1f GB per GPU)" % (2 * 8.0 * s * s / world / 1e9)}


class Harness:
    """Process set-up shared by the workloads: device, stream, context, torch.distributed, watchdog, timing."""

    def __init__(self, args):
        import torch
        import chan_vese_b200 as cv
        self.torch, self.cv, self.args = torch, cv, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.t_start = time.time()
        if args.watchdog > 0:
            # a rank that never comes back (a peer lost, a collective stuck) must not hold the box until the caller's own
            # limit: report and leave -- os._exit tears the CUDA context down with the process
            def give_up():
                if self.rank == 0:
                    print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": self.world,
                                      "error": "watchdog: no result after %d s" % args.watchdog}), flush=True)
                os._exit(3)
            wd = threading.Timer(args.watchdog, give_up)
            wd.daemon = True
            wd.start()
        self.stream = torch.cuda.Stream()
        self.ctx = cv.Context(self.local, stream=self.stream.cuda_stream)

    def trace(self, msg):  # CVB_BENCH_TRACE=1: one stderr line per phase and rank (where does a multi-rank run stop?)
        if os.environ.get("CVB_BENCH_TRACE"):
            print("[rank %d %.1fs] %s" % (self.rank, time.time() - self.t_start, msg), file=sys.stderr, flush=True)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, sample_clocks):
        """W untimed warm-up calls, then exactly `steps` calls between a barrier + synchronize on both sides; device time
        from CUDA events on the launching stream, max over ranks."""
        torch = self.torch
        for k in range(warmup):
            fn()
            self.trace("%s: warm-up step %d done" % (fn.__name__, k))
        self.barrier()
        self.ctx.reset_stats()
        sampler = ClockSampler(self.local) if sample_clocks and self.rank == 0 else None  # one nvidia-smi poller per box
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(self.stream)
        self.trace("%s: %d timed steps enqueued and returned" % (fn.__name__, steps))
        self.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if self.dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, self.ctx.stats(), clocks

    def gather(self, obj):
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        self.ctx.close()
        if self.dist is not None:
            self.dist.destroy_process_group()


# ---- the other BASELINE configurations, timed in the same run (rank 0 of a single-GPU run) ------------------------
def time_config(h, name, fp32=False, reps=3):
    import numpy as np
    cv, torch = h.cv, h.torch
    from chan_vese_b200 import synth
    c = synth.CONFIGS[name]
    hh, ww, n = c["h"], c["w"], c["n"]
    k = dict(c["csv"])
    max_steps = k.pop("max_steps")
    tol = k.pop("tol", 1e-3)
    prm = cv.make_params(nch=n, **k)
    img = {"C1": synth.seastar, "C2": synth.night_lights, "C3": synth.two_phase}[name]()
    sess = cv.Session(h.ctx, n, hh, ww, fp32=fp32)
    sess.upload_image(img)
    sess.save_image()
    u0 = cv.levelset_circ(hh, ww, ww // 2, hh // 2, hh // 4) if c["init"] == "circ" else None

    def run():
        sess.restore_image()
        if u0 is None:
            sess.init_checkerboard()
        else:
            sess.upload_levelset(u0)
        npm = sess.perona_malik(**c["pm"]) if c["pm"] else 0
        steps, _ = sess.csv_run(prm, tol=tol, max_steps=max_steps)
        return npm, steps

    run()
    torch.cuda.synchronize()
    h.ctx.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(h.stream)
    for _ in range(reps):
        npm, steps = run()
    e1.record(h.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = h.ctx.stats()
    peak, _ = measured_peak()
    esz = 4 if fp32 else 8
    csv_us = st["csv_ms"] / reps / max(steps, 1) * 1e3
    pm_us = st["pm_ms"] / reps / max(npm, 1) * 1e3
    res = {"h": hh, "w": ww, "channels": n, "pm_steps": npm, "csv_steps": int(steps), "ms": ms,
           "pixel_iters_per_s": float(hh) * ww * (npm + steps) / (ms * 1e-3),
           "csv_us_per_step": csv_us, "pm_us_per_step": pm_us if npm else None,
           "csv_frac_hbm": (2 * esz + n) * hh * ww / (csv_us * 1e-6) / 1e9 / peak if steps else None,
           "pm_frac_hbm": 2 * esz * n * hh * ww / (pm_us * 1e-6) / 1e9 / peak if npm else None,
           "dtype": "f32" if fp32 else "f64"}
    sess.close()
    return res


def time_batch_sample(h, count=512):
    """configs[4] on this GPU alone: `count` of the 4096 images (what one rank of an 8-GPU run holds), one batch job."""
    import numpy as np
    cv, torch = h.cv, h.torch
    from chan_vese_b200 import synth
    c = synth.CONFIGS["C5"]
    hh, ww, n = c["h"], c["w"], c["n"]
    k = dict(c["csv"])
    max_steps = k.pop("max_steps")
    tol = k.pop("tol", 1e-3)
    prm = cv.make_params(nch=n, **k)
    base = synth.batch_images(0, 64, hh, ww)
    imgs = np.ascontiguousarray(base[np.arange(count) % 64])
    job = cv.Batch(h.ctx, count, n, hh, ww)
    job.upload_images(imgs)
    job.save_images()

    def run():
        job.restore_images()
        job.init_checkerboard()
        npm = job.perona_malik(**c["pm"])
        steps, _ = job.csv_run(prm, tol=tol, max_steps=max_steps)
        return npm, steps

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(h.stream)
    reps = 2
    for _ in range(reps):
        npm, steps = run()
    e1.record(h.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pixit = float(hh) * ww * (npm * count + int(steps.sum()))
    peak, _ = measured_peak()
    alg = (BYTES_CSV_RGB * int(steps.sum()) + BYTES_PM_RGB * npm * count) * hh * ww
    res = {"images": count, "h": hh, "w": ww, "pm_steps": npm, "csv_steps_min": int(steps.min()), "csv_steps_max": int(steps.max()),
           "csv_steps_mean": float(steps.mean()), "ms": ms, "pixel_iters_per_s": pixit / (ms * 1e-3),
           "frac_hbm": alg / (ms * 1e-3) / 1e9 / peak,
           "what": "%d of the 4096 images of configs[4] as one batch job on this GPU (the share of one rank of 8); "
                   "the whole batch on N GPUs: bench.py --workload batch" % count}
    job.close()
    return res


def time_oneshot(h, views):
    """cvb_segment: one call, pageable host buffers, fp64 u in and out, PM planes and the byte mask out."""
    import numpy as np
    cv = h.cv
    s = h.args.size
    planes = [np.array(v) for v in views]  # pageable copies
    u0 = cv.levelset_checkerboard(s, s)
    prm = cv.make_params()
    out = {}
    for it in range(2):  # the first call allocates the context's one-shot session, the second reuses it
        h.ctx.reset_stats()
        t0 = time.perf_counter()
        r = h.ctx.segment(planes, u0, prm, tol=0.0, max_steps=CSV_STEPS, smooth=True, **PM)
        dt = time.perf_counter() - t0
        st = h.ctx.stats()
        assert r["steps"] == CSV_STEPS
        out["first_call_ms" if it == 0 else "ms"] = dt * 1e3
    out["value"] = float(s) * s * (20 + CSV_STEPS) / (out["ms"] * 1e-3)
    out["unit"] = UNIT
    out["h2d_bytes"] = int(st["h2d_bytes"])
    out["d2h_bytes"] = int(st["d2h_bytes"])
    out["what"] = "cvb_segment (second call: buffers cached in the context), pageable numpy buffers, wall clock"
    h.ctx.trim()
    return out


def run_slabs(args):
    import numpy as np
    h = Harness(args)
    cv, torch, ctx, rank, world, dist = h.cv, h.torch, h.ctx, h.rank, h.world, h.dist
    from chan_vese_b200 import synth
    H = W = args.size
    tile_rows = args.tile_rows or cv.auto_tile_rows(H, W, 1, world)
    ctx.set_tile_rows(tile_rows)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(ctx.comm_create_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), world, rank)
        lo, hi = cv.slab_partition(H, tile_rows, world, rank)
        sess = cv.Session(ctx, 3, H, W, rows=(lo, hi))
    else:
        lo, hi = 0, H
        sess = cv.Session(ctx, 3, H, W)
    rows = hi - lo
    # synthetic input in pinned host memory (identical bytes whatever the slab decomposition)
    pinned = [torch.empty((rows, W), dtype=torch.uint8).pin_memory() for _ in range(3)]
    views = [p.numpy() for p in pinned]
    synth.hashed_scene_rows(H, W, lo, hi, out=views, threads=min(8, os.cpu_count() or 1))
    mask_pin = torch.empty((rows, (W + 7) // 8), dtype=torch.uint8).pin_memory()
    mask_view = mask_pin.numpy()
    prm = cv.make_params()
    h.trace("session created, slab rows [%d, %d), tile rows %d" % (lo, hi, tile_rows))
    sess.upload_image(views)
    sess.save_image()
    h.trace("image uploaded")
    last = {}

    def step_resident():
        sess.restore_image()
        sess.init_checkerboard()
        n_pm = sess.perona_malik(PM["K"], PM["L"], PM["T"])
        n_csv, norm = sess.csv_run(prm, tol=0.0, max_steps=CSV_STEPS)
        last.update(n_pm=n_pm, n_csv=n_csv, norm=norm)
        return n_pm, n_csv

    def step_e2e():
        # host buffers in, host buffer out, every step: a stream of images through the session.  restore_image makes the
        # image prefetched during the previous step current; the upload of the NEXT step's planes (pinned host memory ->
        # HBM, on the copy stream) then runs behind this step's diffusion and level-set loop; the segmentation mask is read
        # back bit-packed.  Every step uploads all of its input and downloads its result inside the timed region.
        sess.restore_image()
        sess.prefetch_image(views)
        sess.init_checkerboard()
        n_pm = sess.perona_malik(PM["K"], PM["L"], PM["T"])
        n_csv, norm = sess.csv_run(prm, tol=0.0, max_steps=CSV_STEPS)
        sess.mask_packed(out=mask_view)
        last.update(n_pm=n_pm, n_csv=n_csv, norm=norm)
        return n_pm, n_csv

    ms, (n_pm, n_csv), st, clocks = h.timed(step_resident, args.steps, args.warmup, True)
    assert n_pm == 20 and n_csv == CSV_STEPS, (n_pm, n_csv)
    # ---- what the timed steps computed: the session still holds the last step's result
    parts = [crc_of(p) for p in sess.download_image()] + [crc_of(sess.download_levelset()), crc_of(sess.mask_packed())]
    digest_resident = (last["n_csv"], float(last["norm"]).hex())
    pix_iters = float(H) * W * (n_pm + n_csv) * args.steps
    value = pix_iters / (ms * 1e-3)
    sess.prefetch_image(views)  # the first e2e step's image (the warm-up steps keep the pipeline primed)
    ms_e2e, _, st_e2e, _ = h.timed(step_e2e, args.steps, max(1, args.warmup // 3), False)
    e2e_value = pix_iters / (ms_e2e * 1e-3)
    # the end-to-end sequence must have produced the same mask (its buffer is the pinned one the timed steps filled)
    parts.append(crc_of(mask_view))
    all_parts = h.gather(parts)
    all_digest = h.gather((digest_resident, (last["n_csv"], float(last["norm"]).hex())))
    all_wait = h.gather((st.get("peer_wait_ms", 0.0), st.get("peer_waits", 0), st["csv_ms"], st["csv_step_launches"]))
    if rank != 0:
        sess.close()
        h.close()
        return
    crcs = combine_ranks(all_parts)
    digest = {"steps_done": digest_resident[0], "last_norm_hex": digest_resident[1],
              "pm_planes_crc32": ["%08x" % c for c in crcs[:3]], "levelset_crc32": "%08x" % crcs[3],
              "mask_packed_crc32": "%08x" % crcs[4], "e2e_mask_packed_crc32": "%08x" % crcs[5],
              "ranks_agree": all(d == all_digest[0] for d in all_digest) and all_digest[0][0] == all_digest[0][1],
              "tile_rows": tile_rows,
              "what": "state after the last timed step: CRC32 (zlib) of the uint8 PM planes, of the fp64 level set and of the "
                      "bit-packed mask, each over the whole image in row order (per-rank CRCs combined); identical for every N"}
    peak, peak_src = measured_peak()
    csv_launch_ms = st["csv_ms"] / max(st["csv_step_launches"], 1)
    pm_launch_ms = st["pm_ms"] / max(st["pm_step_launches"], 1)
    pm_steps_per_launch = (n_pm * args.steps) / max(st["pm_step_launches"], 1)  # 2 with temporal blocking
    csv_gbs = BYTES_CSV_RGB * rows * W / (csv_launch_ms * 1e-3) / 1e9
    pm_gbs = BYTES_PM_RGB * pm_steps_per_launch * rows * W / (pm_launch_ms * 1e-3) / 1e9
    alg_bytes_step = (BYTES_CSV_RGB * n_csv + BYTES_PM_RGB * n_pm) * rows * W
    bpp, traffic_src = ncu_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": st_e2e["h2d_bytes"] // args.steps, "d2h_bytes_per_step": st_e2e["d2h_bytes"] // args.steps},
        "gpu_launches": int(st["kernel_launches"]),
        "result_digest": digest,
        "roofline": {"bound": "hbm", "kernel": "csv_step_kernel<3,fast>", "achieved": csv_gbs, "peak": peak, "unit": "GB/s",
                     "frac": csv_gbs / peak, "traffic": bpp * rows * W if bpp else None,
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_CSV_RGB * rows * W, "launch_ms": csv_launch_ms,
                     "share_of_step": st["csv_ms"] / ms},
        "kernels": {"csv_step": {"launches": int(st["csv_step_launches"]), "ms_per_launch": csv_launch_ms, "GBps": csv_gbs,
                                 "frac": csv_gbs / peak, "pixel_iters_per_s": rows * W / (csv_launch_ms * 1e-3)},
                    "pm_step": {"launches": int(st["pm_step_launches"]), "steps_per_launch": pm_steps_per_launch,
                                "ms_per_launch": pm_launch_ms, "ms_per_step_equivalent": pm_launch_ms / pm_steps_per_launch,
                                "GBps": pm_gbs, "frac": pm_gbs / peak,
                                "pixel_iters_per_s": pm_steps_per_launch * rows * W / (pm_launch_ms * 1e-3)},
                    "per_rank": [{"csv_ms_per_launch": a[2] / max(a[3], 1), "peer_wait_us_per_step": 1e3 * a[0] / max(a[1], 1)}
                                 for a in all_wait] if world > 1 else None,
                    "whole_step_GBps": alg_bytes_step * args.steps / (ms * 1e-3) / 1e9,
                    "whole_step_frac": alg_bytes_step * args.steps / (ms * 1e-3) / 1e9 / peak,
                    "note": "value's step also holds an 805 MB device-to-device restore of the image and the device "
                            "checkerboard; pm_step GBps counts the ALGORITHMIC 48 B per pixel and diffusion step (a launch "
                            "that fuses two steps can exceed the HBM peak by this definition, SURVEY 8d)"},
    }
    if world == 1 and not args.no_extra:
        extra = {}
        sess.release_scratch()
        ctx.set_tile_rows(0)  # every configuration with the library's own tile choice
        for name, fp32 in (("C1", False), ("C2", False), ("C3", False), ("C3", True)):
            try:
                extra[name + ("_fp32" if fp32 else "")] = time_config(h, name, fp32=fp32, reps=2 if name == "C3" else 5)
            except Exception as e:  # an extra must never cost the headline line
                extra[name + ("_fp32" if fp32 else "")] = {"error": repr(e)}
        try:
            extra["C5_512"] = time_batch_sample(h)
        except Exception as e:
            extra["C5_512"] = {"error": repr(e)}
        try:
            extra["e2e_oneshot"] = time_oneshot(h, views)
        except Exception as e:
            extra["e2e_oneshot"] = {"error": repr(e)}
        line["extra"] = extra
    if world == 1 and not args.no_cpu:
        v, desc, cores, _ = cpu_sample(args.cpu_size, 1, 0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
    print(json.dumps(line), flush=True)
    sess.close()
    h.close()


def run_batch(args):
    """configs[4]: a batch of independent images, split over the ranks, no communicator."""
    import numpy as np
    h = Harness(args)
    cv, torch, ctx, rank, world = h.cv, h.torch, h.ctx, h.rank, h.world
    from chan_vese_b200 import synth
    c = synth.CONFIGS["C5"]
    hh, ww, n = c["h"], c["w"], c["n"]
    k = dict(c["csv"])
    max_steps = k.pop("max_steps")
    tol = k.pop("tol", 1e-3)
    prm = cv.make_params(nch=n, **k)
    if args.batch % world:
        raise SystemExit("--batch must be a multiple of the rank count")
    count = args.batch // world
    first = rank * count
    # 64 distinct scenes, repeated: generating 4096 of them on the host would take longer than the whole bench
    base = synth.batch_images(0, 64, hh, ww)
    idx = (np.arange(first, first + count) % 64)
    pinned = torch.empty((count, n, hh, ww), dtype=torch.uint8).pin_memory()
    imgs = pinned.numpy()
    imgs[:] = base[idx]
    job = cv.Batch(ctx, count, n, hh, ww)
    job.upload_images(imgs)
    job.save_images()
    last = {}

    def step_resident():
        job.restore_images()
        job.init_checkerboard()
        npm = job.perona_malik(**c["pm"])
        steps, norm = job.csv_run(prm, tol=tol, max_steps=max_steps)
        last.update(npm=npm, steps=steps, norm=norm)
        return npm, steps

    masks = np.empty((count, hh, (ww + 7) // 8), dtype=np.uint8)

    def step_e2e():
        npm = job.upload_images_smooth(imgs, **c["pm"])
        job.init_checkerboard()
        steps, norm = job.csv_run(prm, tol=tol, max_steps=max_steps)
        job.masks_packed(out=masks)
        last.update(npm=npm, steps=steps, norm=norm)
        return npm, steps

    ms, (npm, steps), st, clocks = h.timed(step_resident, args.steps, args.warmup, True)
    pixit_rank = float(hh) * ww * (npm * count + int(steps.sum()))
    steps_crc = crc_of(steps.astype(np.int32))
    norm_crc = crc_of(last["norm"])
    mask_crc = crc_of(job.masks_packed())
    ms_e2e, _, st_e2e, _ = h.timed(step_e2e, args.steps, max(1, args.warmup // 3), False)
    allp = h.gather(([steps_crc, norm_crc, mask_crc, crc_of(masks)], pixit_rank, int(steps.min()), int(steps.max()), int(steps.sum())))
    if rank != 0:
        job.close()
        h.close()
        return
    crcs = combine_ranks([p[0] for p in allp])
    pix_iters = sum(p[1] for p in allp) * args.steps
    total_steps = sum(p[4] for p in allp)
    peak, peak_src = measured_peak()
    alg = (BYTES_CSV_RGB * total_steps + BYTES_PM_RGB * npm * args.batch) * hh * ww * args.steps
    line = {
        "metric": METRIC, "value": pix_iters / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
        "e2e": {"value": pix_iters / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": st_e2e["h2d_bytes"] // args.steps, "d2h_bytes_per_step": st_e2e["d2h_bytes"] // args.steps},
        "gpu_launches": int(st["kernel_launches"]),
        "result_digest": {"csv_steps_min": min(p[2] for p in allp), "csv_steps_max": max(p[3] for p in allp),
                          "csv_steps_total": total_steps, "steps_crc32": "%08x" % crcs[0], "norms_crc32": "%08x" % crcs[1],
                          "masks_packed_crc32": "%08x" % crcs[2], "e2e_masks_packed_crc32": "%08x" % crcs[3],
                          "what": "per-image step counts, last norms and bit-packed masks of the whole batch in image order; "
                                  "identical for every N"},
        "roofline": {"bound": "hbm", "kernel": "whole batch step (csv_step + pm_step launches)", "achieved": alg / (ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak / world, "traffic": None,
                     "peak_source": peak_src + " x n_gpus"},
    }
    print(json.dumps(line), flush=True)
    job.close()
    h.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="slabs", choices=["slabs", "batch"])
    ap.add_argument("--size", type=int, default=16384, help="image side (default: the BASELINE 16384)")
    ap.add_argument("--batch", type=int, default=4096, help="images of the batch workload (whole job, all ranks)")
    ap.add_argument("--cpu-size", type=int, default=1024, help="side of the CPU-baseline crop")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1/C2/C3/one-shot extras")
    ap.add_argument("--tile-rows", type=int, default=0, help="rows per tile (0 = the library's automatic choice)")
    ap.add_argument("--watchdog", type=int, default=420, help="give up after this many seconds (0 = never)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "batch":
        run_batch(args)
    else:
        run_slabs(args)


if __name__ == "__main__":
    main()
