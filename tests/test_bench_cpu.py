"""CPU checks of bench.py: the reference arm prints exactly one JSON line with the contract's keys,
and the algorithmic-byte model matches a hand count on the layer table."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None


def test_reference_arm_non_zero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_hbm_byte_model():
    sys.path.insert(0, ROOT)
    import bench

    class D:      # one 3x3 stride-1 conv layer 16 -> 16 on 2 x 8 x 8
        n, xd, xh, xw, cx = 2, 1, 8, 8, 16
        yd, yh, yw, cy = 1, 8, 8, 16
    e = 2 * 8 * 8 * 16
    want = (e * 2 + e * 4) + e * 4 + (e * 4 + e * 4) + (2 * (e * 2 + e * 4) + e * 2) + (e * 2 + e * 2) + (e * 2 + e * 2)
    assert bench.hbm_bytes([("some/layer", D, False)]) == want
    assert bench.hbm_bytes([("conv0", D, False)]) == want - (e * 2 + e * 2)      # no input gradient for the stem
