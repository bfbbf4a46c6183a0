"""Normalisation: per-row fc[i]/mx.nd.norm(fc[i]) (final_efm.py:240-243,
extract_feacture_v2.py:70,97) and whole-block anc/mx.nd.norm(anc) (train_efm.py:241), with backward."""
from __future__ import annotations

import torch

from . import _lib as L


def _as2d(x):
    if x.dim() == 1:
        return x.unsqueeze(0), True
    return x.reshape(x.shape[0], -1), False


class _L2Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        L.require_cuda(x)
        x2 = x.contiguous()
        B, D = x2.shape
        y = torch.empty_like(x2)
        with torch.cuda.device(x.device):
            L.check(L.lib.efm_l2norm_fwd(L.ptr(x2), L.ptr(y), None, B, D, L.dtype_code(x2), L.stream()))
        ctx.save_for_backward(x2)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        B, D = x.shape
        with torch.cuda.device(x.device):
            L.check(L.lib.efm_l2norm_bwd(L.ptr(x), L.ptr(dy), L.ptr(dx), B, D, L.dtype_code(x), L.stream()))
        return dx


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """Per-row x / ||x||_2 over all non-batch dims (no epsilon: a zero row gives NaN like the reference)."""
    x2, squeeze = _as2d(x)
    y = _L2Norm.apply(x2)
    return y.reshape(x.shape) if not squeeze else y.reshape(x.shape)


def row_norms(x: torch.Tensor) -> torch.Tensor:
    """||x_i||_2 per row as fp32 (the `sqrt_simdot` stored beside every gallery row, Feature.hpp:476)."""
    L.require_cuda(x)
    x2 = x.reshape(x.shape[0], -1).contiguous()
    out = torch.empty(x2.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.lib.efm_l2norm_fwd(L.ptr(x2), None, L.ptr(out), x2.shape[0], x2.shape[1], L.dtype_code(x2), L.stream()))
    return out


class _Frob(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        L.require_cuda(x)
        xc = x.contiguous()
        ss = torch.zeros(1, dtype=torch.float32, device=x.device)
        y = torch.empty_like(xc)
        n = xc.numel()
        with torch.cuda.device(x.device):
            L.check(L.lib.efm_sumsq(L.ptr(xc), n, L.dtype_code(xc), L.ptr(ss), L.stream()))
            L.check(L.lib.efm_frob_fwd(L.ptr(xc), L.ptr(ss), L.ptr(y), n, L.dtype_code(xc), L.stream()))
        ctx.save_for_backward(xc, ss)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, ss = ctx.saved_tensors
        dy = dy.contiguous()
        dot = torch.zeros(1, dtype=torch.float32, device=xc.device)
        dx = torch.empty_like(xc)
        n = xc.numel()
        with torch.cuda.device(xc.device):
            L.check(L.lib.efm_dot(L.ptr(xc), L.ptr(dy), n, L.dtype_code(xc), L.ptr(dot), L.stream()))
            L.check(L.lib.efm_frob_bwd(L.ptr(xc), L.ptr(dy), L.ptr(ss), L.ptr(dot), L.ptr(dx), n, L.dtype_code(xc), L.stream()))
        return dx


def norm(x: torch.Tensor) -> torch.Tensor:
    """mx.nd.norm(x) with no axis: sqrt of the sum of squares of ALL elements, shape (1,), fp32."""
    L.require_cuda(x)
    xc = x.contiguous()
    ss = torch.zeros(1, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.lib.efm_sumsq(L.ptr(xc), xc.numel(), L.dtype_code(xc), L.ptr(ss), L.stream()))
    return ss.sqrt_()


def frob_normalize(x: torch.Tensor) -> torch.Tensor:
    """x / mx.nd.norm(x): the whole-block scaling of train_efm.py:241 (gradient flows through the norm)."""
    return _Frob.apply(x)
