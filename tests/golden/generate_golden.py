"""Regenerates the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python tests/golden/generate_golden.py
It compiles the reference's scheduler/mult/history sources where they lie (oracle/Makefile ->
oracle/_ref/libsgcn_ref.so, oracle/_ref/test_mult), drives them through the same call sequences
as the reference's own smoke programs (gcn/test_scheduler.py, gcn/test_mult.cpp) plus a few seeded
random graphs, and stores every output array.  float32 arrays are stored as uint32 bit patterns so
that the comparison in tests/ is bit-exact.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import native  # noqa: E402
from tests.graphs_small import tree11, random_graph  # noqa: E402

VEC_NAMES = ["field", "ffield", "edg_s", "edg_t", "fedg_s", "fedg_t", "scales", "edg_w", "medg_w", "fedg_w"]


def enc(a):
    a = np.asarray(a)
    if a.dtype == np.float32:
        return {"f32bits": a.view(np.uint32).tolist()}
    return {"i32": a.astype(np.int64).tolist()}


def run_case(adj, seed, degrees, batches, cv, importance):
    """degrees[k] = degree of the k-th expand after start_batch (i.e. already in call order)."""
    s = native.RefSampler(adj.data, adj.indices, adj.indptr, cv=cv, importance=importance, L=len(degrees))
    s.seed(seed)
    out = []
    for ids in batches:
        s.start_batch(np.asarray(ids, dtype=np.int32))
        levels = []
        for d in degrees:
            s.expand(int(d))
            snap = s.snapshot()
            levels.append({k: enc(snap[k]) for k in VEC_NAMES})
        out.append(levels)
    final_adj = {"adj_i": enc(s.vec("adj_i")), "adj_w": enc(s.vec("adj_w"))}
    return {"seed": seed, "degrees": list(map(int, degrees)), "batches": [list(map(int, b)) for b in batches],
            "cv": bool(cv), "importance": bool(importance), "results": out, "final": final_adj}


def main():
    native.build()
    assert native.have_ref(), "reference sources not available: cannot regenerate golden vectors"
    cases = {}
    t = tree11()
    # G1: gcn/test_scheduler.py -- PyScheduler(adj, labels, 2, [1, 2], ..., seed 0, cv=True), batch [0];
    # expand order is degrees[L-l-1] => 2 then 1
    cases["G1_tree_cv_seed0"] = dict(graph="tree11", **run_case(t, 0, [2, 1], [[0]], True, False))
    # G2: NS, seed 1, degrees [1,1], batch [0,1] twice (state persists across batches)
    cases["G2_tree_ns_seed1"] = dict(graph="tree11", **run_case(t, 1, [1, 1], [[0, 1], [0, 1]], False, False))
    # G3: importance sampling, same inputs
    cases["G3_tree_is_seed1"] = dict(graph="tree11", **run_case(t, 1, [1, 1], [[0, 1], [0, 1]], False, True))
    # seeded random graphs (power-law-ish degrees, some empty rows, duplicate-free)
    for gi, (n, avg, gseed) in enumerate([(64, 4, 11), (300, 9, 12), (1000, 20, 13)]):
        g = random_graph(n, avg, gseed)
        rng = np.random.RandomState(100 + gi)
        batches = [rng.choice(n, size=min(n, 17 + 13 * gi), replace=False).tolist() for _ in range(3)]
        tag = "random:%d:%d:%d" % (n, avg, gseed)
        cases["R%d_ns" % gi] = dict(graph=tag, **run_case(g, 5 + gi, [2, 3], batches, False, False))
        cases["R%d_cv" % gi] = dict(graph=tag, **run_case(g, 7 + gi, [1, 2], batches, True, False))
        cases["R%d_exact" % gi] = dict(graph=tag, **run_case(g, 9 + gi, [10000], batches, True, False))
        cases["R%d_is" % gi] = dict(graph=tag, **run_case(g, 3 + gi, [2, 2], batches, False, True))
    with open(os.path.join(HERE, "sampler_golden.json"), "w") as f:
        json.dump(cases, f)

    # G0: gcn/test_mult.cpp stdout
    out = subprocess.run([native.REF_TEST_MULT], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(HERE, "test_mult_stdout.txt"), "w") as f:
        f.write(out)

    # Mult draws + slicers on seeded inputs
    extra = {}
    rng = np.random.RandomState(42)
    for n in (1, 2, 5, 8, 9, 100, 1000):
        p = (rng.rand(n).astype(np.float32) ** 4 + np.float32(1e-6)).astype(np.float32)
        m = native.RefMult(p)
        extra["mult_n%d" % n] = {"prob": enc(p), "bit": enc(m.bit()), "draws": [m.query() for _ in range(n)]}
    g = random_graph(200, 7, 21)
    rows = rng.choice(200, size=37, replace=True).astype(np.int32)
    idx, val, shape = native.ref_slice(g, rows)
    dense = rng.randn(200, 13).astype(np.float32)
    extra["slice"] = {"graph": "random:200:7:21", "rows": rows.tolist(), "idx": enc(idx.reshape(-1)),
                      "val": enc(val), "shape": list(map(int, shape))}
    extra["dense_slice"] = {"seed": 42, "rows": rows.tolist(), "src": enc(dense.reshape(-1)),
                            "out": enc(native.ref_dense_slice(dense, rows).reshape(-1)), "cols": 13}
    with open(os.path.join(HERE, "mult_slice_golden.json"), "w") as f:
        json.dump(extra, f)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
