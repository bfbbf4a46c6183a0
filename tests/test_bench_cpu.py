"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the required keys, and under a
multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_json_line():
    lines = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"])
    d = json.loads(lines[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "embeddings/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("C2") and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_multi_rank_only_rank0_prints():
    assert _run(["--impl", "reference", "--gpus", "8", "--steps", "1", "--warmup", "0"], {"RANK": "3", "WORLD_SIZE": "8"}) == []
    d = json.loads(_run(["--impl", "reference", "--gpus", "8", "--steps", "1", "--warmup", "0"],
                        {"RANK": "0", "WORLD_SIZE": "8", "OMP_NUM_THREADS": "1"})[-1])
    assert d["n_gpus"] == 8 and d["scaling"] == "strong" and "65536" in d["config"]["workload"]
    if (os.cpu_count() or 1) > 1:
        assert d["cpu_baseline"]["cores"] > 1                             # torchrun's OMP_NUM_THREADS=1 is lifted
