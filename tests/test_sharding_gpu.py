"""GPU (needs >= 2 devices, skipped otherwise): the sharded pass under torchrun, both transports,
eager and CUDA-graph replay, against the multi-process oracle of tests/mgpu_check.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("transport,mode,graph", [("nccl", "cv", ""), ("peer", "cv", ""), ("peer", "cvd", "graph"),
                                                  ("nccl", "cvd", "graph"), ("peer", "cv", "pipelined")])
def test_two_rank_pass_matches_oracle(transport, mode, graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py"),
           transport, mode] + ([graph] if graph else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "mgpu_check ok" in out.stdout


@pytest.mark.xfail(reason="multi-GPU form of the gather-ahead schedule: written after the round-1 GPU budget was spent, "
                          "not yet run on hardware", strict=False)
@pytest.mark.parametrize("form", ["ahead", "ahead-graph"])
def test_two_rank_gather_ahead_matches_oracle(form):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import signal
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tests", "mgpu_check.py"),
           "peer", "cv", form]
    # a process group of its own: on a timeout every rank goes, not only the launcher
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT,
                            start_new_session=True)
    try:
        stdout, stderr = proc.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        proc.communicate()
        raise
    assert proc.returncode == 0 and "mgpu_check ok" in stdout, stdout[-3000:] + stderr[-3000:]
