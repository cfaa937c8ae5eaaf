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


def test_clock_sampler_drops_samples_before_mark():
    """The sampler is started before the warm-up steps and mark()ed in front of the timed region: only samples taken after
    the mark count, throttle reasons included."""
    import time

    sys.path.insert(0, ROOT)
    import bench

    class _Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    s = bench.ClockSampler(0)
    s.proc = _Done()
    now = time.time()
    s.lines = [(now - 10.0, "0, 1200, 1965, 400.0, Active, Not Active, Not Active, Not Active"),   # warm-up: dropped
               (now + 1.0, "0, 1650, 1965, 990.0, Not Active, Not Active, Not Active, Active"),
               (now + 1.2, "0, 1670, 1965, 995.0, Not Active, Not Active, Not Active, Active"),
               (now + 1.4, "garbage")]
    s.t0 = 0.0
    allc = s.stop()
    assert allc["samples"] == 3 and "hw_slowdown" in allc["reasons"]
    s.t0 = now
    c = s.stop()
    assert c["samples"] == 2 and c["sm_mhz"] == 1660.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
