"""bench.py's output contract, checked on the CPU: the reference arm prints the JSON line the driver parses, and the
product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _run(args, timeout=300):
    return subprocess.run([sys.executable, BENCH] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("workload,metric", [("c2", "nerf_train_samples_per_s"), ("c1", "mlp_fit_train_samples_per_s")])
def test_reference_arm_prints_the_contract_line(workload, metric):
    out = _run(["--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0"])
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == metric and line["unit"] == "samples/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_on_other_ranks_prints_nothing():
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=ROOT, capture_output=True, text=True,
                         timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA")
    out = _run(["--steps", "1", "--warmup", "0"], timeout=120)
    assert out.returncode != 0
    assert "CUDA" in (out.stderr + out.stdout)


def test_timed_region_is_cut_into_whole_rotations_of_the_batch_pool():
    """bench.py replays CUDA graphs of consecutive steps; every graph but a last remainder must be a multiple of the batch
    pool, so each replay walks the whole rotation from batch 0 and no batch is revisited while it can still sit in L2."""
    import types
    sys.path.insert(0, ROOT)
    import bench
    for n_pool in (2, 6, 7):
        for T in (1, 2, 5, 6, 7, 50, 240, 241, 1100, 5000):
            obj = types.SimpleNamespace(n_pool=n_pool, CHUNK=bench.TrainRun.CHUNK)
            plan = bench.TrainRun.chunk_plan(obj, T)
            assert sum(plan) == T and all(c > 0 for c in plan)
            if T > n_pool:
                assert all(c % n_pool == 0 for c in plan[:-1]) and max(plan) <= bench.TrainRun.CHUNK
                assert len(set(plan[:-1])) <= 1          # one graph for the full chunks, one for the remainder
