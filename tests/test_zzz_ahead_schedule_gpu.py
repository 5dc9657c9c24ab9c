"""GPU: the "gather ahead" schedule (csrc/step.cu:sgcn_step_run_ahead, HotPathStep.run_ahead / capture_ahead /
replay_ahead) vs the CPU oracle: eager and captured as CUDA graphs, CV / CVD / NS, both normalisations,
batches that share nodes.  The check runs in a process of its own (tests/ahead_check.py, with a timeout).
Status: parity-green on a B200 at the very end of round 1 (gpurun_out/ahead/log.txt); not timed yet --
bench.py offers it as --driver ahead, the default driver is unchanged until it has been measured."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gather_ahead_schedule_matches_oracle():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ahead_check.py")], capture_output=True,
                         text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0 and "ahead_check ok" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]


def test_gather_ahead_host_buffer_forms_match_oracle():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ahead_host_check.py")], capture_output=True,
                         text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0 and "ahead_host_check ok" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]
