"""bench.py contract pieces that can be checked without a GPU: the reference arm (the reference's arithmetic on the
host cores) prints ONE JSON line with the agreed keys, also under torchrun where only rank 0 speaks; the product arm
refuses to run without CUDA instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _json_lines(text):
    return [json.loads(line) for line in text.splitlines() if line.startswith("{")]


def test_reference_arm_line():
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True,
                         text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "sweep cell*angle*group updates/s" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None and d["value"] > 0
    assert d["config"]["workload"].startswith("synthetic 70-group slab, S256, 1M cells")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sweep sets" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_under_torchrun_only_rank0_prints():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", BENCH, "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1", "--no-cpu-baseline"], capture_output=True,
                         text=True, cwd=ROOT, timeout=300)
    assert out.returncode != 0 and not _json_lines(out.stdout)      # no number without the CUDA path
