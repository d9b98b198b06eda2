"""bench.py's CPU arm (`--impl reference`) prints ONE JSON line with the keys the driver reads; the
arm runs the reference's own NumPy for S1/S2 where it can be imported. Two frames on two processes
(no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-workers", "2"], capture_output=True, text=True, check=True,
                         cwd=ROOT, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["metric"] == "front-end frames/sec"
    cb = d["cpu_baseline"]
    assert cb["cores"] == 2 and cb["kind"] == "port" and cb["value"] == d["value"]
    assert set(cb["kind_by_stage"]) == {"S1", "S2", "S3", "S4", "S5"}
    from oracle import ref_shim
    assert cb["kind_by_stage"]["S1"] == ("reference" if ref_shim.available() else "port")
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    assert set(cb["stage_ms_per_frame"]) >= {"S1", "S2", "S4"}
