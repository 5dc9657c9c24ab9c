"""CPU: bench.py's host logic -- chunk-length choice, committed-capture lookup, and the reference arm
(the reference's CPU path: compiled reference sampler + slicer, C port of the aggregate) end to end
on a shrunken graph, checking the JSON contract of the line it prints."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_graph_length_choice():
    b = _bench()
    for k in list(range(2, 70)) + [100, 200, 1000, 4992, 5000]:
        s = b.pick_graph_passes(k, 64)
        assert 2 <= s <= 64
        if k <= 64:
            assert s == k                                      # short runs: ONE graph of exactly K passes
    assert b.pick_graph_passes(20, 64) == 20 and b.pick_graph_passes(5000, 64) == 50
    assert b.pick_graph_passes(2048, 64) == 64 and b.pick_graph_passes(131, 64) == 64   # prime: remainder passes


def test_both_arms_print_the_same_config():
    b = _bench()
    import argparse
    a = argparse.Namespace(workload="reddit_cv", scale=1.0, seed=1)
    c = b.workload_config(a, b.WORKLOADS["reddit_cv"])
    assert set(c) == {"workload", "scale", "seed", "l2"} and "model" not in c


def test_traffic_comes_from_the_committed_capture():
    b = _bench()
    t = b.load_traffic("full_mean_kernel")
    assert t is not None and t["bytes"] > 1e6 and t["source"].startswith("profiles/")
    assert b.load_traffic("no_such_kernel") is None
    peak, src = b.load_peaks()
    assert 3000 < peak < 9000 and src


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.02",
                          "--steps", "3", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                     # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "edges/s" and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
