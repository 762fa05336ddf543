"""bench.py's reference arm runs without a GPU: its JSON line must carry the keys of the measurement contract
(metric / unit / config of BASELINE.json, cpu_baseline describing the run, e2e repeating the value with zero transfers)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-size", "128"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["unit"] == "pixel-iterations/s" and line["metric"] in base["metric"]
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "gpu_launches"):
        assert k in line, k
    assert line["value"] > 0 and line["dtype"] == "f64" and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0  # nothing of the product ran on this arm


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--cpu-size", "128"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_result_digest_crc_combination():
    """bench.py combines per-rank CRC32s into the CRC32 of the whole image in row order (zlib's crc32_combine): the
    digest of a slab run must equal the digest of the single-GPU run of the same bytes, however the rows are split."""
    import zlib

    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, size=(97, 53), dtype=np.uint8)
    whole = zlib.crc32(data.tobytes())
    for cuts in ([0, 97], [0, 40, 97], [0, 1, 2, 50, 96, 97], [0, 13, 26, 39, 52, 65, 78, 91, 97]):
        parts = [[bench.crc_of(data[a:b])] for a, b in zip(cuts, cuts[1:])]
        assert bench.combine_ranks(parts) == [whole]
    assert bench.crc32_combine(zlib.crc32(b"abc"), zlib.crc32(b""), 0) == zlib.crc32(b"abc")
    u = rng.standard_normal((64, 31))
    assert bench.combine_ranks([[bench.crc_of(u[:20])], [bench.crc_of(u[20:])]]) == [zlib.crc32(u.tobytes())]


def test_bench_line_keys_are_documented_in_the_source():
    """The JSON line of the GPU arm cannot be produced here (no GPU); its contract keys must at least be built in bench.py."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"', '"scaling"',
                '"vs_baseline"', '"dtype"', '"data"', '"config"', '"clocks"', '"e2e"', '"h2d_bytes_per_step"', '"d2h_bytes_per_step"',
                '"gpu_launches"', '"roofline"', '"traffic"', '"cpu_baseline"', '"result_digest"'):
        assert key in src, key
