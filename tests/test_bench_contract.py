"""bench.py's reference arm runs on the CPU, so its JSON contract can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_ours_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "tiny", "--steps", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)


def test_ncu_traffic_table_covers_the_bench_plans():
    """bench.py fills roofline.traffic from profiles/ncu_traffic.json, keyed by the plan name the library reports
    (skm_lloyd_kernel_name); the plans of the headline (config 2) and of the north-star shape (config 3) must be there."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "ncu_traffic.json")) as f:
        table = json.load(f)
    for name in ("k_assign_fast64<10>", "k_prefix16<64> x1 on 4 of the entry pairs + k_assign_bounded"):
        assert name in table and table[name]["dram_bytes_per_point"] > 0 and "source" in table[name]
    # algorithmic bytes per point (8 m + 8): the pruned plan's traffic stays within 1.3x of it
    assert table["k_prefix16<64> x1 on 4 of the entry pairs + k_assign_bounded"]["dram_bytes_per_point"] <= 1.3 * (51 * 8 + 8)
