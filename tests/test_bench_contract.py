"""bench.py's reference arm runs on the host cores alone (the oracle port; no JVM exists here), so its
JSON line can be checked on CPU: the keys the driver reads, and the same metric / config as the B200 arm.
The B200 arm must refuse to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--streams", "64", "--enc-steps", "1", "--ref-enc-blocks", "8",
             "--c5-blocks", "32")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MB/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench
    assert d["metric"] == bench.METRIC
    # the reference arm also times the encoder (C3 blocks) and the sharded-corpus shape (C5), all host cores
    assert d["encode"]["value"] > 0 and d["encode"]["kind"] == "port" and d["encode"]["cores"] == d["cpu_baseline"]["cores"]
    assert d["c5"]["value"] > 0 and d["c5"]["decode"]["value"] > 0


def test_cpu_sample_is_class_balanced():
    import bench
    for n, want in [(2048, 256), (2048, 4096), (64, 8), (2048, 2048)]:
        pick = bench.cpu_sample_blocks(n, want)
        assert len(set(pick.tolist())) == len(pick) and pick.max() < n
        counts = [int((pick % 4 == c).sum()) for c in range(4)]
        assert len(set(counts)) == 1


def test_reference_arm_other_ranks_stay_silent():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--streams", "8", "--no-encode", "--no-c5", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_refuses_to_run_without_a_gpu(lzb):
    if lzb.lib().lzb_device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--no-encode", "--no-c5", "--no-cpu", "--streams", "8")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "no CPU fallback" in (r.stderr + r.stdout)
