"""bench.py --impl reference (the compiled reference's CPU path, what the driver times next to the GPU arm) must survive the
driver's own --steps / --warmup: reduced-size warm-up frames are followed by full-size ones, and the reference indexes
previousFrameGrid with the NEW frame's pixels (render_utils.cpp:154), so a predecessor of another resolution must not be
handed over as history (it was, with --warmup >= 2: a segmentation fault and an empty reference line).  CPU only."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libromis_ref.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libromis_ref.so not built (make -C oracle ref)")
@pytest.mark.parametrize("steps,warmup", [(2, 3), (1, 1)])
def test_reference_arm_prints_one_line_for_any_warmup(steps, warmup):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1",
                        "--steps", str(steps), "--warmup", str(warmup)], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, f"rc {r.returncode}\n{r.stderr[-2000:]}"
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == steps and d["warmup"] == warmup and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
