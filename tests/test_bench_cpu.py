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


def test_chunk_length_choice():
    b = _bench()
    for k in list(range(1, 70)) + [100, 200, 1000, 4992, 5000]:
        s = b.pick_steps_per_graph(k, 16)
        assert s % 2 == 0 and 2 <= s <= 16
        cost = lambda c: (k // c) * 8 + (k % c) * 40
        assert all(cost(s) <= cost(c) for c in range(2, min(16, max(k, 2)) + 1, 2))
    assert b.pick_steps_per_graph(5000, 16) == 16 and b.pick_steps_per_graph(20, 16) == 10
    assert b.pick_steps_per_graph(64, 8) == 8


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
