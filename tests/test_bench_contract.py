"""bench.py's contract, as far as it can be checked without a GPU: the reference arm runs, prints ONE JSON line with
the keys the driver reads, and carries exactly the config our arm prints for the same workload (`same_config`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("workload", ["cfg2", "cfg3"])
def test_reference_arm_prints_the_contract_line(workload):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] == 0
    # the same config object our arm prints for this workload: the driver compares them
    assert d["config"] == bench.workload_config(workload, 1, None)
    assert "cpu_arm_sample" in d["config"]


def test_both_arms_share_config_and_method_override_is_marked():
    a = bench.workload_config("cfg2", 8, "ETD35")
    b = bench.workload_config("cfg2", 8, None)
    assert a == b and "x8" in a["parallelism"]
    c = bench.workload_config("cfg2", 1, "IF45DP")
    assert c["method"] == "IF45DP" and "overridden" in c["workload"]


def test_ncu_inventory_feeds_the_traffic_figure():
    """roofline.traffic comes from the committed ncu inventory by kernel name, not from a literal."""
    traffic, co = bench.traffic_of("nl_fast_pre_kernel", "cfg2")
    assert traffic is not None and 0.9e9 < traffic < 1.2e9          # 537 MB read + 480-ish MB written per launch
    assert co["source"].startswith("profiles/") and 0 < co["fp64_pipe_pct"] < 100
    assert bench.traffic_of("no_such_kernel", "cfg2") == (None, None)


def test_cpu_arm_runs_every_workload_sample():
    for workload, rows, method, size in (("cfg2", 2, "ETD35", None), ("cfg3", 4, "ETD4", None), ("cfg4", 0, "IF45DP", 64),
                                         ("cfg5", 0, "ETD35", 16)):
        value, wall, work, kind = bench.cpu_baseline(workload, 1, 1, 0, rows, method, 0.0, size)
        assert value > 0 and work > 0 and kind in ("reference", "port")
