"""bench.py's contract on a machine without a GPU: the reference arm (the unmodified reference binary on the host
cores) prints one JSON line with the agreed keys; the product arm refuses to run rather than fall back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line(built):
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-liks", "2e5", "--config", "cfg3_staph"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "read-placement likelihoods/sec" and d["unit"] == "likelihoods/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    # the config object is the workload's identity, the same keys the product arm prints
    assert d["config"]["workload"] == "cfg3_staph" and d["config"]["gaps"] == 100
    assert set(d["config"]) == {"workload", "seed", "gaps", "scaling", "partition", "l2"}
    assert "seeded random subset" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]


def test_product_arm_needs_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--config", "cfg3_staph"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_reference_arm_under_torchrun_prints_once(built):
    """Launched like the product arm for N > 1: rank 0 alone runs the reference and prints; the other ranks exit 0."""
    if not refrun.have_ref():
        pytest.skip("oracle/_ref binaries not built")
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--ref-liks", "2e5", "--config", "cfg3_staph"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference" and json.loads(lines[0])["n_gpus"] == 2
