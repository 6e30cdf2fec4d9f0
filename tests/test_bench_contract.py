"""bench.py --impl reference runs on the host cores only: check its one-JSON-line contract without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "segmented points/sec (fwd)" and d["unit"] == "points/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0
    assert d["n_gpus"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert list(d["config"]) == ["workload"] and d["config"]["workload"].startswith("configs[0]")
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_both_arms_name_the_same_config_and_the_package_never_imports_the_oracle():
    """The driver compares the two arms' `config` dicts: they must be equal for every workload. And the product package holds
    no import of oracle/ (the CPU legs live in bench.py / bench_nn.py at the repo root)."""
    import re
    src = open(os.path.join(ROOT, "bench_nn.py")).read() + open(os.path.join(ROOT, "bench.py")).read()
    ref = open(os.path.join(ROOT, "oracle", "nn_bench.py")).read() + open(os.path.join(ROOT, "bench.py")).read()
    for tag in ("configs[0]: segmentation forward, batch", "configs[2]: training step fwd+loss+bwd+2xAdam, batch",
                "configs[1]: FPS %d windows x %d pts -> %d per GPU, float32 rows of %d columns", "k-means assignment pass, %d points x 3 features, k = %d\""):
        assert tag in src and tag in ref, tag
    pkg = os.path.join(ROOT, "3d-semantic-segmentation-amp-net_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            text = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), name
