"""bench.py's contract on a machine without a GPU: the reference arm (the CPU port of the
reference's NumPy path, the only leg that may execute ``oracle/``) prints exactly one JSON line
with the agreed keys, ranks other than 0 stay silent under torchrun, and the GPU arm fails loudly
instead of falling back to anything."""

import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=600):
    full_env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        full_env.pop(k, None)
    full_env.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=full_env,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_agreed_keys():
    res = _run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--cpu-per-core", "4"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    line = json.loads(lines[0])
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        baseline = json.load(fh)
    assert baseline["metric"].startswith(line["metric"])
    assert line["impl"] == "reference" and line["unit"] == "propagations/s"
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["config"]["workload"].startswith("configs[1]") and "65536 instances per GPU" in line["config"]["workload"]
    assert "model" not in line["config"]
    cores = os.cpu_count()
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] == cores and cpu["unit"] == line["unit"]
    assert "%d instances per step" % (4 * cores) in cpu["sample"] and "oracle/ref_fixed.py" in cpu["sample"]
    assert cpu["value"] == line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
    # value = instances per step / time per step
    assert abs(line["value"] - 4 * cores / (line["ms_per_step"] / 1e3)) <= 1e-6 * line["value"]


def test_reference_arm_is_silent_on_the_other_ranks():
    res = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
               env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0, res.stderr[-2000:]
    assert res.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    res = _run(["--steps", "1", "--warmup", "1", "--configs", "none", "--skip-cpu"])
    assert res.returncode != 0
    assert res.stdout.strip() == ""                       # no JSON line: nothing was measured


def test_only_the_cpu_legs_of_bench_touch_the_oracle():
    """``oracle/`` is test infrastructure: in bench.py it may be imported by the CPU baseline /
    reference arm (``_cpu_work``) and by the checker of the all-gather leg, nowhere else; the
    timed hot path (jt_bench_lib.HotPath) never."""
    src = open(BENCH).read()
    owners = []
    current = None
    for text in src.splitlines():
        m = re.match(r"def (\w+)\(", text)
        if m:
            current = m.group(1)
        if re.search(r"^\s*(from oracle|import oracle)", text):
            owners.append(current)
    assert sorted(owners) == ["_cpu_work", "gather_leg"], owners
    lib = open(os.path.join(ROOT, "junction-tree_b200", "jt_bench_lib.py")).read()
    assert not re.search(r"^\s*(from oracle|import oracle)", lib, flags=re.M)
