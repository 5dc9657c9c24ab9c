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
