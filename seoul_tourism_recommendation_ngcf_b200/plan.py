"""Device-resident execution plan for one ``lap_list`` element.

The reference keeps each Laplacian as an uncoalesced int64 ``torch.sparse_coo`` (matrix.py:79-83) and lets
``torch.mm`` re-sort it on every call (NGCF.py:130; again, transposed, in the backward).  The plan converts it
ONCE into int32 CSR for L and for L^T, with the permutations back to COO order so a per-edge node-dropout
mask generated in COO order serves both directions, plus the hub-row split tables the SpMM kernel uses.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class CsrSide:
    """CSR of L (or of L^T) on the device: rowptr/colidx/perm int32, base values fp32, hub split tables."""

    def __init__(self, rowptr, colidx, perm, vals, n_rows, chunk):
        self.rowptr, self.colidx, self.perm, self.vals = rowptr, colidx, perm, vals
        self.n_rows = int(n_rows)
        lib = _lib.load()
        split = lib.ngcf_spmm_split_threshold()
        deg = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
        hub = torch.nonzero(deg > split).flatten()
        self.n_hub = int(hub.numel())
        dev = rowptr.device
        if self.n_hub:
            hdeg = deg[hub]
            nch = (hdeg + chunk - 1) // chunk
            cptr = torch.zeros(self.n_hub + 1, dtype=torch.int64, device=dev)
            cptr[1:] = torch.cumsum(nch, 0)
            n_chunks = int(cptr[-1])
            owner = torch.repeat_interleave(torch.arange(self.n_hub, device=dev), nch)
            local = torch.arange(n_chunks, device=dev) - cptr[owner]
            beg = rowptr[hub][owner].to(torch.int64) + local * chunk
            end = torch.minimum(beg + chunk, rowptr[hub + 1][owner].to(torch.int64))
            self.hub_rows = hub.to(torch.int32)
            self.hub_chunk_ptr = cptr.to(torch.int32)
            self.hub_chunk_begin = beg.to(torch.int32)
            self.hub_chunk_end = end.to(torch.int32)
            self.hub_chunk_row = hub[owner].to(torch.int32)
            self.n_chunks = n_chunks
        else:
            self.hub_rows = self.hub_chunk_ptr = self.hub_chunk_begin = self.hub_chunk_end = self.hub_chunk_row = None
            self.n_chunks = 0
        self._partial = {}

    def hub_partial(self, d: int):
        if not self.n_hub:
            return None
        buf = self._partial.get(d)
        if buf is None:
            buf = torch.empty(self.n_chunks * d, dtype=torch.float32, device=self.rowptr.device)
            self._partial[d] = buf
        return buf


class LaplacianPlan:
    def __init__(self, L: torch.Tensor, device, hub_chunk: int = 256):
        if not (L.is_sparse and L.dim() == 2 and L.shape[0] == L.shape[1]):
            raise ValueError("lap_list entries must be square torch.sparse_coo tensors (matrix.py:79-83)")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("LaplacianPlan needs a CUDA device: the NGCF B200 path has no CPU fallback")
        lib = _lib.load()
        self.src = L
        self.N = int(L.shape[0])
        idx = L._indices().to(device=device, dtype=torch.int64).contiguous()
        self.coo_val = L._values().to(device=device, dtype=torch.float32).contiguous()
        self.nnz = int(self.coo_val.numel())
        if self.nnz and (int(idx.min()) < 0 or int(idx.max()) >= self.N):
            raise ValueError("Laplacian indices out of range")
        row, col = idx[0].contiguous(), idx[1].contiguous()
        need = C.c_size_t(0)
        _lib.check(lib.ngcf_coo_to_csr_workspace(self.nnz, self.N, C.byref(need)), "coo_to_csr_workspace")
        ws = torch.empty(need.value, dtype=torch.uint8, device=device)
        sides = []
        for transpose in (0, 1):
            rowptr = torch.empty(self.N + 1, dtype=torch.int32, device=device)
            colidx = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=device)
            perm = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=device)
            _lib.check(lib.ngcf_coo_to_csr(row.data_ptr(), col.data_ptr(), self.nnz, self.N, self.N, transpose,
                                           rowptr.data_ptr(), colidx.data_ptr(), perm.data_ptr(), ws.data_ptr(),
                                           need.value, _stream()), "coo_to_csr")
            vals = torch.empty(max(self.nnz, 1), dtype=torch.float32, device=device)
            _lib.check(lib.ngcf_edge_values(self.coo_val.data_ptr(), perm.data_ptr(), None, vals.data_ptr(), self.nnz,
                                            _stream()), "edge_values")
            sides.append(CsrSide(rowptr, colidx, perm, vals, self.N, hub_chunk))
        torch.cuda.current_stream().synchronize()
        del ws
        self.fwd, self.bwd = sides
        # L = D^-1/2 A D^-1/2 is symmetric (matrix.py:48-62): then L^T's CSR is L's, and sharing the arrays
        # halves the distinct bytes a training step touches.  Node dropout breaks the symmetry per step, which
        # is handled by giving the two directions separate masked value arrays.
        self.symmetric = bool(torch.equal(self.fwd.rowptr, self.bwd.rowptr) and torch.equal(self.fwd.colidx, self.bwd.colidx)
                              and torch.equal(self.fwd.vals, self.bwd.vals))

    def side(self, transposed: bool, masked: bool) -> CsrSide:
        if transposed and not (self.symmetric and not masked):
            return self.bwd
        return self.fwd

    def masked_values(self, side: CsrSide, keep_mask):
        """CSR-ordered edge values with an explicit COO-order node-dropout mask folded in (NGCF.py:93-100)."""
        lib = _lib.load()
        out = torch.empty_like(side.vals)
        _lib.check(lib.ngcf_edge_values(self.coo_val.data_ptr(), side.perm.data_ptr(), _lib.ptr(keep_mask),
                                        out.data_ptr(), self.nnz, _stream()), "edge_values")
        return out


def spmm(side: CsrSide, vals, X, d: int, out=None, addend=None, slot=None, gsum=None, drop_p: float = 0.0,
         seed: int = 0, seed_dev=None, layer: int = 0, transposed: bool = False):
    """Y = L·X (+ addend) (+ gsum[slot] rows) through ngcf_spmm; drop_p > 0 = in-kernel device-RNG node dropout."""
    lib = _lib.load()
    if out is None:
        out = torch.empty(side.n_rows, d, dtype=torch.float32, device=X.device)
    _lib.check(lib.ngcf_spmm(side.rowptr.data_ptr(), side.colidx.data_ptr(), vals.data_ptr(), side.n_rows,
                             X.data_ptr(), X.stride(0), d,
                             _lib.ptr(addend), addend.stride(0) if addend is not None else 0,
                             _lib.ptr(slot), _lib.ptr(gsum), gsum.stride(0) if gsum is not None else 0,
                             _lib.ptr(side.hub_rows), _lib.ptr(side.hub_chunk_ptr), side.n_hub,
                             _lib.ptr(side.hub_chunk_begin), _lib.ptr(side.hub_chunk_end), _lib.ptr(side.hub_chunk_row),
                             side.n_chunks, _lib.ptr(side.hub_partial(d)),
                             float(drop_p), int(seed) & (2 ** 64 - 1), _lib.ptr(seed_dev), int(layer), int(transposed),
                             out.data_ptr(), out.stride(0), _stream()), "spmm")
    return out
