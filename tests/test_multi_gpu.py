"""Two (or more) GPUs of one box: every rank runs its domain of ONE world (CUDA path, halo exchange over peer memory
or NCCL); the decomposed run must reproduce the single-domain oracle run -- same infections per step, same final
person state (tests/mgpu_check.py under torchrun).  Skipped on a box with a single GPU."""
import os
import subprocess
import sys

import pytest

from helpers import have_gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("halo", ["p2p", "nccl", "p2p-serial-arrival"])
def test_decomposed_run_equals_single_domain_oracle(halo):
    assert have_gpu()
    n = n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    ranks = 2 if n < 4 else 4
    env = dict(os.environ)
    env.pop("JB_HALO", None)
    if halo == "nccl":
        env["JB_HALO"] = "nccl"
    if halo == "p2p-serial-arrival":
        env["JB_HALO_ASYNC"] = "0"  # the visitors' arrival in front of every interaction kernel (round-1 order)
    port = {"p2p": 29541, "nccl": 29542, "p2p-serial-arrival": 29543}[halo]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mgpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    line = [l for l in out.splitlines() if l.startswith("mgpu_check:")]
    log = os.environ.get("JB_MGPU_LOG")
    if log:
        with open(log, "a") as f:
            f.write(f"[{halo}, {ranks} ranks] " + (line[-1] if line else out[-2000:]) + "\n")
    assert r.returncode == 0 and line and line[-1].endswith("OK"), out[-3000:]
