"""The bench contract's reference arm (`bench.py --impl reference`) runs on the host cores only: its JSON line can be checked on
the CPU tier.  (The GPU arm of bench.py needs a B200 and is exercised by the driver / `gpurun`.)"""
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def _run(*extra):
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--objects", "3000", "--steps", "2", "--warmup", "1", *extra],
                         capture_output=True, text=True, timeout=600, cwd=str(REPO))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run()
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "objects/s"
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert abs(d["ms_per_step"] * d["value"] / 1e3 - 3000) < 1e-3 * 3000  # value = objects / time of one step
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"]
    assert str(cb["cores"]) in cb["scaling"]                      # reported at a thread count of the ladder ...
    best = max(cb["scaling"].values())
    assert cb["scaling"][str(cb["cores"])] >= 0.5 * best          # ... the one the ladder pass found fastest (timing noise aside)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["same_config"] is True and d["config"]["sample_objects"] == 3000
    assert cb["literal_port"]["value"] > 0
