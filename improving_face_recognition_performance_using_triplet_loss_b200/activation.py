"""EFM-3 / MFM-2 max-out activation.

Drop-in for the MXNet op sequence SliceChannel(num_outputs=3, axis=1) -> maximum, maximum, minimum,
minimum -> Concat of the reference (lightcnn.py:22-27,32-37,53-58,61-66,123-128; efm_symbol.py:63-77),
one fused CUDA pass forward and one backward.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L


def _act_shape(x: torch.Tensor, mode: int):
    if x.dim() < 2:
        raise ValueError("expected (N, C, ...) input, got shape %s" % (tuple(x.shape),))
    N, Cc = x.shape[0], x.shape[1]
    HW = 1
    for s in x.shape[2:]:
        HW *= s
    if Cc % mode:
        raise ValueError("channel count %d is not divisible by %d (SliceChannel would fail)" % (Cc, mode))
    out_c = Cc // 3 * 2 if mode == 3 else Cc // 2
    return N, Cc, HW, out_c


class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode):
        L.require_cuda(x)
        x = x.contiguous()
        N, Cc, HW, out_c = _act_shape(x, mode)
        y = torch.empty((N, out_c) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.lib.efm_act_fwd(mode, L.ptr(x), L.ptr(y), N, Cc, HW, L.dtype_code(x), L.stream()))
        ctx.save_for_backward(x)
        ctx.mode = mode
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        mode = ctx.mode
        dy = dy.contiguous()
        N, Cc, HW, _ = _act_shape(x, mode)
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(L.lib.efm_act_bwd(mode, L.ptr(x), L.ptr(dy), L.ptr(dx), N, Cc, HW, L.dtype_code(x), L.stream()))
        return dx, None


def efm3(x: torch.Tensor) -> torch.Tensor:
    """concat(max(s0,s1,s2), min(s0,s1,s2)) over the three contiguous channel thirds: C -> 2C/3."""
    return _Act.apply(x, L.ACT_EFM3)


def mfm2(x: torch.Tensor) -> torch.Tensor:
    """max(s0, s1) over the two channel halves: C -> C/2 (the classic LightCNN MFM)."""
    return _Act.apply(x, L.ACT_MFM2)


class EFM(nn.Module):
    """The activation alone (what the reference inlines after every conv / dense)."""

    def __init__(self, num_outputs: int = 3):
        super().__init__()
        if num_outputs not in (2, 3):
            raise ValueError("num_outputs must be 3 (EFM) or 2 (MFM)")
        self.num_outputs = num_outputs

    def forward(self, x):
        return efm3(x) if self.num_outputs == 3 else mfm2(x)


class efm(nn.Module):
    """Mirror of the gluon block `efm(num_filter, num_filter1, kernel_size, stride, padding, efm_type)`
    (lightcnn.py:6-39): efm_type 1 = 1x1 conv -> EFM -> conv -> EFM, otherwise conv -> EFM.  The convolutions
    stay in the host framework (cuDNN); only the max-out is ours."""

    def __init__(self, in_channels, num_filter, num_filter1, kernel_size, stride, padding, efm_type):
        super().__init__()
        self.efm_type = efm_type
        c = in_channels
        if efm_type == 1:
            self.conv_op_1 = nn.Conv2d(c, num_filter, kernel_size=1)
            c = num_filter // 3 * 2
        self.conv_op_2 = nn.Conv2d(c, num_filter1, kernel_size=kernel_size, stride=stride, padding=padding)

    def forward(self, x):
        if self.efm_type == 1:
            x = efm3(self.conv_op_1(x))
        return efm3(self.conv_op_2(x))
