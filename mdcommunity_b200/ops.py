"""Operator-level drop-in: torch_sparse.spmm(index, value, m, n, matrix) backed by the CSR kernel.

Reference call sites: MultiDismantler_net_graphsage.py:139,146,199,290,297,350,377; MultiDismantler_torch.py:420
(torch-sparse 0.6.18: out = scatter_add(matrix[col] * value[:, None], row, dim_size=m); duplicates are summed).
"""
import ctypes as C

import torch

from . import _lib


def coo_to_csr(index, m):
    """Stable sort by row: keeps the COO order inside each row, so sums run in the reference's order."""
    row, col = index[0], index[1]
    order = torch.argsort(row, stable=True)
    counts = torch.bincount(row, minlength=m)
    rowptr = torch.zeros(m + 1, dtype=torch.int32, device=index.device)
    rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    return rowptr, col[order].to(torch.int32).contiguous(), order


def spmm_csr(rowptr, col, val, matrix, m):
    if not matrix.is_cuda:
        raise RuntimeError("mdcommunity_b200.ops.spmm needs CUDA tensors (no CPU fallback)")
    x = matrix.contiguous().float()
    D = x.shape[1]
    if D % 4:
        raise ValueError("feature width must be a multiple of 4")
    y = torch.empty(m, D, dtype=torch.float32, device=x.device)
    if col.numel() == 0:  # all rows empty: the kernel still needs a non-null column pointer
        col = torch.zeros(1, dtype=torch.int32, device=x.device)
        val = None
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    rc = _lib.lib().md_spmm_csr_f32(C.c_void_p(rowptr.data_ptr()), C.c_void_p(col.data_ptr()),
                                    C.c_void_p(val.data_ptr()) if val is not None else None,
                                    C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), m, D, stream)
    _lib.check(rc, "md_spmm_csr_f32")
    return y


class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, index, value, m, n, matrix):
        rowptr, col, order = coo_to_csr(index, m)
        val = value[order].contiguous().float()
        ctx.save_for_backward(index, value, matrix)
        ctx.m, ctx.n = m, n
        return spmm_csr(rowptr, col, val, matrix, m)

    @staticmethod
    def backward(ctx, grad_out):
        index, value, matrix = ctx.saved_tensors
        g_val = g_mat = None
        if ctx.needs_input_grad[4]:  # d/dX = A^T @ grad
            t_index = torch.stack([index[1], index[0]])
            rowptr, col, order = coo_to_csr(t_index, ctx.n)
            g_mat = spmm_csr(rowptr, col, value[order].contiguous().float(), grad_out.contiguous(), ctx.n)
        if ctx.needs_input_grad[1]:
            g_val = (grad_out[index[0]] * matrix[index[1]]).sum(1)
        return None, g_val, None, None, g_mat


def spmm(index, value, m, n, matrix):
    """Same signature and semantics as torch_sparse.spmm for 2-D `matrix` (participates in autograd)."""
    return _SpmmFn.apply(index, value, m, n, matrix)
