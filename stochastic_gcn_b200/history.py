"""Row slicers -- drop-in for the reference's ``history`` extension (gcn/_history.pyx:25-62).

``slice(a, r)`` and ``dense_slice(a, r)`` keep the reference signatures and return types (NumPy in,
NumPy out) but run on the GPU; ``DeviceCSR`` / ``DeviceDense`` keep the source matrix resident in
HBM so that per-step calls move only the index vector in and the sliced rows out, and the
``*_device`` variants keep the result on the GPU as well.
"""
import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import ops


class DeviceDense:
    """A dense [N, C] float32 matrix resident in HBM (input features, PP features, history)."""

    def __init__(self, a, device=None):
        if isinstance(a, torch.Tensor):
            self.t = a.to(device=device or a.device, dtype=torch.float32).contiguous()
        else:
            self.t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device or "cuda")
        self.shape = tuple(self.t.shape)


class DeviceCSR:
    """A CSR float32/int32 matrix resident in HBM (sparse input features)."""

    def __init__(self, a, device=None):
        device = device or "cuda"
        self.data = torch.from_numpy(np.ascontiguousarray(a.data, dtype=np.float32)).to(device)
        self.indices = torch.from_numpy(np.ascontiguousarray(a.indices, dtype=np.int32)).to(device)
        self.indptr = torch.from_numpy(np.ascontiguousarray(a.indptr, dtype=np.int32)).to(device)
        self.shape = tuple(a.shape)
        self.dtype = a.dtype


def _rows_to_device(r, device):
    if isinstance(r, torch.Tensor):
        return r.to(device=device, dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(r, dtype=np.int32)).to(device)


def dense_slice_device(a, r):
    """out[i, :] = a[r[i], :] on the GPU; returns a CUDA tensor (gcn/history.cpp:74-88)."""
    a = a if isinstance(a, DeviceDense) else DeviceDense(a)
    return ops.gather_rows(a.t, _rows_to_device(r, a.t.device))


def dense_slice(a, r):
    """history.dense_slice(a, r) -> float32[len(r), C]  (gcn/_history.pyx:53-62)."""
    return dense_slice_device(a, r).cpu().numpy()


def slice_device(a, r):
    """COO slice of CSR rows on the GPU: (idx2[nnz,2] int32, val[nnz] float32, indptr[n+1] int32)."""
    a = a if isinstance(a, DeviceCSR) else DeviceCSR(a)
    return ops.csr_slice(a.data, a.indices, a.indptr, _rows_to_device(r, a.data.device))


def slice(a, r):
    """history.slice(a, r) -> (int32[nnz,2], float32[nnz], int32[2]); an empty csr_matrix when the
    slice has no stored entry (gcn/_history.pyx:25-51)."""
    n = len(r)
    idx2, val, _ = slice_device(a, r)
    if val.numel() == 0:
        return csr_matrix((n, a.shape[1]), dtype=a.dtype)
    return idx2.cpu().numpy(), val.cpu().numpy(), np.array([n, a.shape[1]], dtype=np.int32)
