"""Multi-GPU parity on real hardware (owner mode was only checked by eye): 2, 4 and 8 ranks under torchrun, each
compares its owned gradient rows and the job-wide energies with the CPU oracle - through the NCCL collectives and through
the exchange over NVLink peer memory (tests/mgpu_worker.py).  Needs >= 2 visible GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`); skipped otherwise."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_owner_mode_under_torchrun_matches_the_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('"ok": true') == 4 * world, r.stdout[-3000:]
