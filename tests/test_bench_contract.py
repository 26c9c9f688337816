"""bench.py contract checks that need no GPU: the reference arm's JSON line, and the no-fallback rule."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "1080p lane-detect frames/s" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_sources_never_touch_the_oracle_or_the_reference_tree():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/; nothing may read /root/reference
    at run time (docstrings citing reference lines are fine)."""
    pkg = os.path.join(ROOT, "multimodal_autonomous_driving_perception_and_planning_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            src = open(os.path.join(base, fn), encoding="utf-8").read()
            assert "import oracle" not in src and "from oracle" not in src, fn
            for line in src.splitlines():
                code = line.split("#", 1)[0].split("//", 1)[0]
                assert "open('/root/reference" not in code and 'open("/root/reference' not in code, (fn, line)
