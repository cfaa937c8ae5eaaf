"""CPU test of the bench.py contract the driver depends on: the reference arm runs without a GPU, prints exactly one JSON
line with the agreed keys, and finishes quickly on a bounded sample."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def _run(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line_with_contract_keys():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "8"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "bdlru_fwd_bwd_seq_tokens_per_s" and d["unit"] == "seq-tokens/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_scoring_workload():
    d = _run(["--impl", "reference", "--workload", "score1m", "--steps", "1", "--warmup", "1", "--cpu-sample", "4"])
    assert d["metric"] == "fullsort_scored_users_per_s" and d["unit"] == "users/s" and d["value"] > 0
