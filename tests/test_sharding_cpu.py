"""CPU: host-side logic of the multi-GPU path -- row-range partition, shard-local CSR, payload
layout, and the exchange protocol itself run by two gloo ranks on CPU tensors (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stochastic_gcn_b200 import _lib
from stochastic_gcn_b200.graphs import powerlaw_graph
from stochastic_gcn_b200.sharding import (HEADER_INTS, merge_payloads, payload_layout, restrict_rows, row_range,
                                          unpack_payload)


def test_row_ranges_partition_the_nodes():
    for n in (1, 7, 232_965, 2_000_000):
        for world in (1, 2, 4, 8):
            spans = [row_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_restricted_csr_keeps_local_rows_and_global_columns():
    g = powerlaw_graph(500, 8000, seed=2, device="cpu", max_degree=100)
    full = g.to_scipy()
    for world in (2, 4):
        total = 0
        for r in range(world):
            lo, hi = row_range(g.n, r, world)
            loc = restrict_rows(g, lo, hi)
            assert loc.indptr.numel() == g.n + 1 and int(loc.indptr[0]) == 0
            m = loc.to_scipy()
            assert (m[lo:hi] != full[lo:hi]).nnz == 0
            assert m[:lo].nnz == 0 and m[hi:].nnz == 0
            total += loc.nnz
        assert total == g.nnz


def test_payload_layout_matches_the_library():
    lib = _lib.load()
    for nb, d in [(0, 4), (1, 1), (1536, 128), (1023, 32), (777, 100)]:
        ids, rows, total = payload_layout(nb, d)
        assert ids == HEADER_INTS * 4 and rows % 16 == 0 and total % 256 == 0
        assert total == lib.sgcn_wb_payload_bytes(nb, d)


def make_payload(ids, rows, n_bound, d):
    ids_off, rows_off, total = payload_layout(n_bound, d)
    buf = np.zeros(total, dtype=np.uint8)
    buf[:4].view(np.int32)[0] = len(ids)
    buf[ids_off:ids_off + 4 * len(ids)].view(np.int32)[:] = ids
    buf[rows_off:rows_off + 4 * rows.size].view(np.float32)[:] = rows.reshape(-1)
    return buf


def _worker(rank, world, port, n_nodes, n_bound, d, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(10 + rank)
    lo, hi = row_range(n_nodes, rank, world)
    hist = np.zeros((n_nodes, d), np.float32)
    for step in range(3):
        # this rank refreshed its batch rows (local range) and some neighbours (anywhere)
        ids = np.unique(np.concatenate([rng.randint(lo, hi, 5), rng.randint(0, n_nodes, 9)])).astype(np.int32)
        rows = (rng.randn(len(ids), d) + 100 * rank + step).astype(np.float32)
        send = torch.from_numpy(make_payload(ids, rows, n_bound, d))
        recv = torch.zeros(world * send.numel(), dtype=torch.uint8)
        dist.all_gather_into_tensor(recv, send)
        slots = recv.numpy().reshape(world, -1)
        merge_payloads(hist, slots, n_bound, d)
    ret[rank] = hist
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_protocol_two_gloo_ranks():
    """Both replicas end identical, and contended rows hold the highest rank's value."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    n_nodes, n_bound, d, world = 40, 16, 8, 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_nodes, n_bound, d, ret), nprocs=world, join=True)
        h0, h1 = ret[0], ret[1]
    assert np.array_equal(h0, h1)
    # replay on one process: rank order, later rank overwrites
    hist = np.zeros((n_nodes, d), np.float32)
    rngs = [np.random.RandomState(10 + r) for r in range(world)]
    for step in range(3):
        for r in range(world):
            lo, hi = row_range(n_nodes, r, world)
            ids = np.unique(np.concatenate([rngs[r].randint(lo, hi, 5), rngs[r].randint(0, n_nodes, 9)])).astype(np.int32)
            rows = (rngs[r].randn(len(ids), d) + 100 * r + step).astype(np.float32)
            hist[ids] = rows
    assert np.array_equal(h0, hist)


def test_unpack_round_trip():
    rng = np.random.RandomState(0)
    ids = rng.permutation(100)[:13].astype(np.int32)
    rows = rng.randn(13, 20).astype(np.float32)
    i2, r2 = unpack_payload(make_payload(ids, rows, 32, 20), 32, 20)
    assert np.array_equal(i2, ids) and np.array_equal(r2, rows)


def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stochastic_gcn_b200 import nn

    class P:                                        # the two fields allreduce_gradients touches
        def __init__(self, shape, seed):
            self.data = torch.zeros(shape, requires_grad=True)
            self.data.grad = torch.from_numpy(np.random.RandomState(seed).randn(*shape).astype(np.float32))

    params = [P((7, 5), 100 + rank), P((5,), 200 + rank), P((3, 2), 300 + rank)]
    if rank == 1:
        params[1].data.grad = None                 # no gradient on this rank: counts as zeros
    nn.allreduce_gradients(params)
    ret[rank] = [p.data.grad.numpy().copy() for p in params]
    dist.barrier()
    dist.destroy_process_group()


def test_dense_gradient_allreduce_two_gloo_ranks():
    """replicated dense weights, gradients averaged over the ranks in one flat all-reduce"""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
        g0, g1 = ret[0], ret[1]
    shapes, seeds = [(7, 5), (5,), (3, 2)], [100, 200, 300]
    for k, (shape, seed) in enumerate(zip(shapes, seeds)):
        a = np.random.RandomState(seed).randn(*shape).astype(np.float32)
        b = np.random.RandomState(seed + 1).randn(*shape).astype(np.float32)
        if k == 1:
            b = np.zeros(shape, np.float32)
        want = (a + b) / 2
        assert np.allclose(g0[k], want, atol=1e-7) and np.array_equal(g0[k], g1[k])


def test_nccl_transport_is_rejected_by_the_multi_pass_drivers():
    """with transport='nccl' the write-back only happens in _finish_exchange; the multi-pass graph
    driver must refuse instead of silently never updating the history."""
    import pytest
    from stochastic_gcn_b200.sharding import ShardedHotPathStep
    s = ShardedHotPathStep.__new__(ShardedHotPathStep)
    s.transport, s.mode = "nccl", "cv"
    with pytest.raises(RuntimeError, match="transport='peer'"):
        s.capture_pipelined(None, None)
    with pytest.raises(RuntimeError, match="transport='peer'"):
        s.run_pipelined([])
    s.mode = "ns"
    s._no_nccl_inside_graphs("x")          # plain sampling keeps no history: nothing to exchange
