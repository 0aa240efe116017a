"""`bench.py --impl reference` (the CPU arm the driver runs next to ours) on a small workload: it must
print one JSON line with the contract's keys and OUR arm's config keys, without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--points", "16384",
         "--steps", "1", "--warmup", "0"],
        capture_output=True, text=True, timeout=600, cwd=ROOT,
        env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 0
    # same config keys as the GPU arm (bench.py: run_ours)
    assert set(line["config"]) == {"workload", "residual_points_per_step", "ic_points", "bc_points",
                                   "parallelism", "l2"}
    assert line["config"]["residual_points_per_step"] == 16384
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert cb["sample_points_per_step"] <= 16384
    assert line["e2e"] == {"value": line["value"], "unit": "points/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}


def test_gpu_arm_config_has_the_same_keys():
    """Static check: the `config` dict literal of run_ours carries exactly those keys too."""
    import ast

    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    configs = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Dict):
            for k, v in zip(node.keys, node.values):
                if isinstance(k, ast.Constant) and k.value == "config" and isinstance(v, ast.Dict):
                    configs.append({kk.value for kk in v.keys if isinstance(kk, ast.Constant)})
    assert len(configs) >= 2
    assert all(c == configs[0] for c in configs), configs
