"""Multi-GPU correctness on real hardware: spawns torchrun with 2 ranks when >= 2 GPUs are visible (skipped otherwise).
The scripts assert internally; a non-zero exit code fails the test and the tail of the output is shown."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(script: str, nproc: int = 2, timeout: int = 900) -> str:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    out = (r.stdout or "") + (r.stderr or "")
    assert r.returncode == 0, f"{script} failed (rc {r.returncode}):\n{out[-6000:]}"
    return out


needs_two = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")


@needs_two
def test_data_parallel_training_paths_agree():
    """NVLink step tail (multicast + peer variants, eager + graphs) vs NCCL all-reduce vs a single process."""
    out = _torchrun("multigpu_train_check.py")
    assert "multi-GPU data-parallel training OK" in out


@needs_two
def test_evaluation_collectives():
    """Parcel-sharded Pearson (all-to-all), statistics all-reduce, rank-aware compute_multidim_pearson, metric classes,
    ensemble averaging, retrieval rank gather — each vs the single-GPU / float64 answer."""
    out = _torchrun("multigpu_eval_check.py")
    assert "multi-GPU evaluation collectives OK" in out
