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
    """SURVEY 8d byte model: no bf16 shadow, no separate statistics pass (fused into the conv epilogue)."""
    sys.path.insert(0, ROOT)
    import bench

    class D:      # one 3x3 stride-1 conv layer 16 -> 16 on 2 x 8 x 8
        n, xd, xh, xw, cx = 2, 1, 8, 8, 16
        yd, yh, yw, cy = 1, 8, 8, 16
    e = 2 * 8 * 8 * 16
    want = (e * 2 + e * 4) + (e * 4 + e * 2) + e * (2 + 4) + e * (2 + 4 + 2) + (e * 2 + e * 2) + (e * 2 + e * 2)
    assert bench.hbm_bytes([("some/layer", D, False)]) == want
    assert bench.hbm_bytes([("conv0", D, False)]) == want - (e * 2 + e * 2)      # no input gradient for the stem
    assert bench.hbm_bytes([("some/layer", D, False)], 4, 4) == (e * 4 + e * 4) + (e * 4 + e * 4) + e * 8 + e * 12 + e * 8 + e * 8
    # mix16 with fp16 raw conv outputs: 2 bytes wherever raw is touched (not for the stem / head)
    assert bench.hbm_bytes([("some/layer", D, False)], 2, 2, 2) == want - 4 * e * 2
    assert bench.hbm_bytes([("conv2", D, False)], 2, 2, 2) == want


def test_launch_groups_pool_equal_shapes_and_pick_the_bound():
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench

    class D16:    # 3x3 16 -> 16 on 32 x 512 x 512: HBM-bound (intensity 48 FLOP/B << ridge)
        n, xd, xh, xw, cx, yd, yh, yw, cy, kd, kh, kw, s = 32, 1, 512, 512, 16, 1, 512, 512, 16, 1, 3, 3, 1

    class D512:   # 3x3 512 -> 512 on 32 x 16 x 16: tensor-bound
        n, xd, xh, xw, cx, yd, yh, yw, cy, kd, kh, kw, s = 32, 1, 16, 16, 512, 1, 16, 16, 512, 1, 3, 3, 1
    table = [("a/b", D16, False), ("c/d", D512, False), ("e/f", D16, False)]
    pk = dict(hbm=6540.8, tf_sus=1373.9, tf_burst=1667.1, src="measured")
    ms = np.zeros((3, 8))
    ms[0, 0], ms[2, 0], ms[1, 4] = 0.6, 0.4, 0.1    # two profiled steps: 0.3 + 0.2 ms fwd of the two 16->16 layers
    groups = bench.launch_groups(table, ms, 2, pk, "2d")
    assert len(groups) == 2
    g = max(groups, key=lambda x: x["ms_per_step"])
    e = 32 * 512 * 512 * 16
    assert g["bound"] == "hbm" and g["unit"] == "GB/s" and g["kernel"] == "conv_fwd 16->16 k3 s1 @512"
    assert g["launches_per_step"] == 2 and g["layers"] == ["a/b", "e/f"] and abs(g["ms_per_launch"] - 0.25) < 1e-12
    assert abs(g["achieved"] - 2 * (e * 2 + e * 4) / 0.5e-3 / 1e9) < 1e-6 and abs(g["frac"] - g["achieved"] / 6540.8) < 1e-12
    assert g["algorithmic_bytes_per_launch"] == e * 6
    assert g["traffic"] == bench.NCU_TRAFFIC[("2d", "16->16 k3 s1 @512", "conv_fwd")]
    t = [x for x in groups if x is not g][0]
    assert t["bound"] == "tensor" and t["unit"] == "TFLOP/s" and t["traffic"] is None
    assert abs(t["achieved"] - 2.0 * 32 * 16 * 16 * 9 * 512 * 512 / 0.05e-3 / 1e12) < 1e-9
