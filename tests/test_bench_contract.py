"""The driver's contract for bench.py, checked on CPU through the reference arm (the only arm that runs without a GPU):
stdout carries exactly ONE line, it is JSON, and it has the keys the contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "c1",
                        "--cpu-windows", "8"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, p.stdout[:500]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "ref k-mers screened/s" and j["unit"] == "kmers/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert j["config"]["workload"].startswith("c1") and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == "kmers/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout) and p.stdout.strip() == ""
