"""bench.py contract pieces that run without a GPU: the reference arm's JSON line
(`--impl reference`, the CPU restatement timed on host cores) and the refusal of the product
arm to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-sample", "20000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "entity_substep_updates_per_sec" and d["unit"] == "entity-substeps/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    # the line names what actually ran: a scaled sample is labelled as one (round 1 printed the 16M label here)
    assert "20000 entities" in d["config"]["workload"] and "scaled sample" in d["config"]["workload"]
    assert d["config"]["entities"] == 20001
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] == d["value"] and "20000 entities" in cb["sample"]
    assert cb["sampled"] is True and cb["sample_entities"] == 20001
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_shortens_a_run_that_exceeds_its_host_time_budget_and_says_so():
    r = run_bench("--impl", "reference", "--steps", "6", "--warmup", "3", "--cpu-sample", "100000", "--ref-budget-s", "0.05")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["steps"] == 6 and d["warmup"] == 3                 # what was asked for
    assert d["steps_run"] == 2 and d["warmup_run"] == 1          # what the budget paid for: never fewer than 1 + 2 frames
    assert "shortened to fit --ref-budget-s" in d["cpu_baseline"]["sample"] and d["value"] > 0
    # within the budget nothing changes
    r = run_bench("--impl", "reference", "--steps", "3", "--warmup", "2", "--cpu-sample", "20000")
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["steps_run"] == 3 and d["warmup_run"] == 2 and "shortened" not in d["cpu_baseline"]["sample"]


def test_reference_arm_on_other_ranks_prints_nothing():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "5000", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and not r.stdout.strip()


def test_product_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
