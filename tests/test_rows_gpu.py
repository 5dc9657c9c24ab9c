"""GPU: row movers and the CSR slicer through the C ABI vs the CPU oracle -- bit-exact (pure copies)."""
import numpy as np
import pytest
import torch

from oracle import native
from tests.conftest import assert_bits_equal, dec, graph_from_tag
from tests.graphs_small import random_graph

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n_src,c,n", [(50, 1, 7), (200, 13, 37), (333, 32, 100), (1000, 128, 513),
                                       (77, 1433, 40), (3000, 1204, 1531), (64, 500, 0)])
def test_dense_slice_bit_exact(n_src, c, n):
    from stochastic_gcn_b200 import ops
    rng = np.random.RandomState(c)
    src = rng.randn(n_src, c).astype(np.float32)
    rows = rng.randint(0, n_src, size=n).astype(np.int32)
    want = native.oracle_dense_slice(src, rows)
    got = ops.gather_rows(dev(src), dev(rows))
    assert_bits_equal(got.cpu().numpy(), want, "gather")
    # strided source / destination (a column block of a wider matrix)
    wide = torch.zeros((n, c + 8), device="cuda")
    ops.gather_rows(dev(src), dev(rows), out=wide[:, 4:4 + c])
    assert_bits_equal(wide[:, 4:4 + c].cpu().numpy(), want, "gather into a strided block")
    assert float(wide[:, :4].abs().sum()) == 0 and float(wide[:, 4 + c:].abs().sum()) == 0
    # device-side row count
    if n > 3:
        out = torch.full((n, c), -1.0, device="cuda")
        ops.gather_rows(dev(src), dev(rows), out=out, n_dev=torch.tensor([n - 3], dtype=torch.int32, device="cuda"))
        assert_bits_equal(out[:n - 3].cpu().numpy(), want[:n - 3], "n_dev")
        assert bool((out[n - 3:] == -1).all())


def test_dense_slice_golden_and_reference_api(mult_slice_golden):
    from stochastic_gcn_b200 import history
    c = mult_slice_golden["dense_slice"]
    src = dec(c["src"]).reshape(-1, c["cols"])
    out = history.dense_slice(src, np.array(c["rows"], dtype=np.int32))
    assert out.dtype == np.float32 and out.shape == (len(c["rows"]), c["cols"])
    assert_bits_equal(out.reshape(-1), dec(c["out"]), "dense_slice golden")


def test_csr_slice_golden_and_oracle(mult_slice_golden):
    from stochastic_gcn_b200 import history
    c = mult_slice_golden["slice"]
    g = graph_from_tag(c["graph"])
    idx, val, shape = history.slice(g, np.array(c["rows"], dtype=np.int32))
    assert idx.dtype == np.int32 and val.dtype == np.float32 and shape.dtype == np.int32
    assert_bits_equal(idx.reshape(-1), dec(c["idx"]), "slice idx")
    assert_bits_equal(val, dec(c["val"]), "slice val")
    assert shape.tolist() == c["shape"]
    # larger ragged case incl. repeated and empty rows, > one scan tile
    g = random_graph(5000, 12, 77)
    rng = np.random.RandomState(1)
    rows = rng.randint(0, 5000, size=4500).astype(np.int32)
    want = native.oracle_slice(g, rows)
    got = history.slice(g, rows)
    assert_bits_equal(got[0], want[0], "idx"); assert_bits_equal(got[1], want[1], "val")
    # all-empty slice returns an empty csr_matrix like the reference
    empty = np.nonzero(np.diff(g.indptr) == 0)[0][:5].astype(np.int32)
    res = history.slice(g, empty)
    assert not isinstance(res, tuple) and res.shape == (len(empty), 5000) and res.nnz == 0


@pytest.mark.parametrize("d", [4, 32, 100, 128, 130, 33])
def test_history_update_and_copy_pad(d):
    from stochastic_gcn_b200 import ops
    rng = np.random.RandomState(d)
    table = rng.randn(500, d).astype(np.float32)
    idx = rng.permutation(500)[:123].astype(np.int32)
    rows = rng.randn(123, d).astype(np.float32)
    want = table.copy()
    native.oracle_lib()  # ensure built
    want[idx] = rows     # tf.scatter_update, unique indices
    t = dev(table)
    ops.history_update(t, dev(idx), dev(rows))
    assert_bits_equal(t.cpu().numpy(), want, "scatter_update")
    out = torch.full((200, d), 7.0, device="cuda")
    ops.copy_rows_pad(dev(rows), 123, out)
    assert_bits_equal(out[:123].cpu().numpy(), rows, "copy")
    assert float(out[123:].abs().sum()) == 0
