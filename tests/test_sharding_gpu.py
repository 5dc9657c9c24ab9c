"""GPU (needs >= 2 devices, skipped otherwise): the sharded pass under torchrun, both transports,
eager and CUDA-graph replay, against the multi-process oracle of tests/mgpu_check.py.  The multi-pass
schedules run on every GPU of the box (world = torch.cuda.device_count()); each run leaves its last lines in
gpurun_out/mgpu_check_<world>.log."""
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


@pytest.mark.parametrize("form,mode,tables", [("trains", "cv", "replicated"), ("trains-graph", "cv", "replicated"),
                                              ("trains", "cvd", "replicated"), ("trains", "cv", "sharded"),
                                              ("trains-graph", "cvd", "sharded")])
def test_all_ranks_multi_pass_schedules_match_oracle(form, mode, tables):
    """the trains schedule (bench default), replicated and sharded tables, on EVERY GPU of the box (2, 4 or 8 ranks)"""
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs 2 GPUs")
    import signal
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tests", "mgpu_check.py"),
           "peer", mode, form, tables]
    # a process group of its own: on a timeout every rank goes, not only the launcher
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT,
                            start_new_session=True)
    try:
        stdout, stderr = proc.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        proc.communicate()
        raise
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "mgpu_check_%d.log" % world), "a") as f:
        f.write("%s %s %s world=%d rc=%d: %s\n" % (form, mode, tables, world, proc.returncode,
                                                    stdout.strip().splitlines()[-1:]))
    assert proc.returncode == 0 and "mgpu_check ok" in stdout, stdout[-3000:] + stderr[-3000:]


@pytest.mark.parametrize("mode,tables", [("cv", "replicated"), ("cvd", "replicated"), ("cvd", "sharded")])
def test_two_ranks_at_the_benched_shapes(mode, tables):
    """BASELINE configs[2] (Reddit-shaped CV+PP degree 2) and configs[3] (CVD+PP degree 1) at FULL size, batch 512,
    width 128, on two ranks: graphs of the trains schedule against R reference samplers sharing one history --
    last pass's rows within 1e-4, every rank's history replica / shard bit-equal to the oracle's table."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import signal
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_check.py"),
           "peer", mode, "trains-graph", tables]
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT,
                            start_new_session=True, env=dict(os.environ, MGPU_FULLSIZE="1"))
    try:
        stdout, stderr = proc.communicate(timeout=420)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        proc.communicate()
        raise
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "mgpu_check_fullsize.log"), "a") as f:
        f.write("%s %s rc=%d: %s\n" % (mode, tables, proc.returncode, stdout.strip().splitlines()[-1:]))
    assert proc.returncode == 0 and "mgpu_check ok" in stdout, stdout[-3000:] + stderr[-3000:]


def test_two_ranks_claims_off_the_chain():
    """the opt-in split form of the ring exchange (SGCN_WB_SPLIT=1: sgcn_wb_claim_ring on the side stream,
    sgcn_wb_copy_ring on the chain) against the same multi-process oracle"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.join(ROOT, "tests", "mgpu_check.py"),
           "peer", "cv", "trains-graph", "replicated"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=dict(os.environ, SGCN_WB_SPLIT="1"))
    assert out.returncode == 0 and "mgpu_check ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
