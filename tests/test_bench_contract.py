"""bench.py's reference arm runs on CPU: its ONE JSON line must carry the keys the driver reads (the GPU arm prints the
same keys plus the roofline; it is exercised on the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

from oracle_tools import have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not have_ref(23), reason="oracle/_ref was not built here")
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--genome", "dm6", "--reads", "2000000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "overlap-scan genome positions/sec" and d["unit"] == "positions/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["steps"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    # like for like with the GPU arm: `value` times the scan + scoring functions only, `e2e` adds the stack build
    # (countReadsInBamFile), so it is the smaller number
    assert 0 < d["e2e"]["value"] < d["value"] and d["e2e"]["value"] == d["cpu_baseline"]["e2e_value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    st = d["cpu_baseline"]["stage_seconds"]
    assert abs(d["ms_per_step"] - 1e3 * (st["t_heights"] + st["t_group"] + st["t_collapse"] + st["t_fdr"])) < 1e-6


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
