"""Triplet-loss head (SURVEY rows a1, a2, a2').

* `TripletLoss` mirrors `gluon.loss.TripletLoss(margin)(pred, positive, negative)` as the reference calls
  it (pre-trained_efm_v3.py:183,210; train_efm.py:210,241; final_efm.py:220,259): per-sample
  relu(sum((p-a)^2 - (n-a)^2) + margin), shape (B,), no mean; backward honours the head gradient and only
  reaches tensors that are in the autograd graph (the reference's negatives are detached copies).
* `sample_negatives` replays the reference's rejection sampling (train_efm.py:234-239,268-273).
* `BatchHardTripletLoss` / `batch_hard_triplet` is the north-star extension: online batch-hard mining fused
  with the loss and its gradient, no B x B matrix in memory.
"""
from __future__ import annotations

import random as _pyrandom

import torch
from torch import nn

from . import _lib as L


# ------------------------------------------------------------------------------------------------ a1
class _Triplet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, p, n, margin):
        L.require_cuda(a, p, n)
        if not (a.shape == p.shape == n.shape):
            raise ValueError("pred/positive/negative shapes differ: %s %s %s" % (a.shape, p.shape, n.shape))
        if not (a.dtype == p.dtype == n.dtype):
            raise ValueError("pred/positive/negative dtypes differ")
        B = a.shape[0]
        a2, p2, n2 = (t.reshape(B, -1).contiguous() for t in (a, p, n))
        D = a2.shape[1]
        loss = torch.empty(B, dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            L.check(L.lib.efm_triplet_fwd(L.ptr(a2), L.ptr(p2), L.ptr(n2), B, D, L.dtype_code(a2), float(margin),
                                          L.ptr(loss), L.stream()))
        ctx.save_for_backward(a2, p2, n2)
        ctx.margin = float(margin)
        ctx.shape = a.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        a2, p2, n2 = ctx.saved_tensors
        B, D = a2.shape
        need = ctx.needs_input_grad
        g = g.contiguous().to(torch.float32)
        da = torch.empty_like(a2) if need[0] else None
        dp = torch.empty_like(a2) if need[1] else None
        dn = torch.empty_like(a2) if need[2] else None
        with torch.cuda.device(a2.device):
            L.check(L.lib.efm_triplet_bwd(L.ptr(a2), L.ptr(p2), L.ptr(n2), L.ptr(g), B, D, L.dtype_code(a2), ctx.margin,
                                          L.ptr(da), L.ptr(dp), L.ptr(dn), L.stream()))
        rs = lambda t: None if t is None else t.reshape(ctx.shape)  # noqa: E731
        return rs(da), rs(dp), rs(dn), None


def triplet_loss(pred, positive, negative, margin=1.0):
    """Functional form: (B,) fp32 per-sample losses."""
    return _Triplet.apply(pred, positive, negative, margin)


class TripletLoss(nn.Module):
    """gluon.loss.TripletLoss(margin=1, weight=None, batch_axis=0) on torch tensors."""

    def __init__(self, margin=1, weight=None, batch_axis=0, **kwargs):
        super().__init__()
        if batch_axis != 0:
            raise ValueError("batch_axis must be 0 (the reference never uses another)")
        self._margin = margin
        self._weight = weight
        self._batch_axis = batch_axis

    def forward(self, pred, positive, negative, sample_weight=None):
        loss = triplet_loss(pred, positive, negative, self._margin)
        if sample_weight is not None:                       # gluon _apply_weighting
            loss = loss * sample_weight.reshape(loss.shape).to(loss.dtype)
        if self._weight is not None:
            loss = loss * self._weight
        return loss

    hybrid_forward = forward


# ------------------------------------------------------------------------------------------------ a2
def sample_negatives(labels, batch_size, pool="train", rng=None):
    """The reference's negative selection, draw for draw: j = randint(0, hi) until label[j] != label[i];
    hi = B-1 for training, 2B-1 for validation/test.  Labels are read to the host ONCE (the reference pays two
    device syncs per draw); returns a LongTensor of row indices on the labels' device."""
    rng = rng or _pyrandom
    hi = batch_size - 1 if pool == "train" else 2 * batch_size - 1
    lab = [int(float(v)) for v in torch.as_tensor(labels).reshape(-1).tolist()]
    if len(lab) <= hi:
        raise ValueError("labels has %d entries, pool needs %d" % (len(lab), hi + 1))
    if len(set(lab[: hi + 1])) < 2:
        raise ValueError("single-identity pool: the reference's rejection loop would never terminate")
    out = []
    for i in range(batch_size):
        j = rng.randint(0, hi)
        while lab[j] == lab[i]:
            j = rng.randint(0, hi)
        out.append(j)
    dev = labels.device if isinstance(labels, torch.Tensor) else "cpu"
    return torch.tensor(out, dtype=torch.long, device=dev)


# ------------------------------------------------------------------------------------------------ a2'
def _labels_i32(labels, device):
    lab = torch.as_tensor(labels, device=device).reshape(-1)
    if lab.dtype != torch.int32:
        lab = lab.to(torch.int32)                           # float labels (MXNet iterators) are exact integers
    return lab.contiguous()


def _bh_flags(normalize, detach_gathered, mine_bf16):
    return (L.BH_NORMALIZE if normalize else 0) | (L.BH_DETACH_GATHERED if detach_gathered else 0) | \
           (L.BH_MINE_BF16 if mine_bf16 else 0)


def batch_hard_triplet(E_all, labels_all, margin, *, normalize=True, row_offset=0, num_local=None, head_grad=None,
                       want_grad=True, detach_gathered=False, mine_bf16=None, workspace=None):
    """Fused batch-hard triplet loss forward (+ backward) -- one call into libefmtl.

    E_all (B_all, D) is the (all-gathered) batch; anchors are rows [row_offset, row_offset+num_local).
    Returns (loss (B,), idx_p (B,), idx_n (B,), dE_all (B_all, D) fp32 or None).
    mine_bf16: None = tensor-core mining for bf16/fp16 inputs, exact fp32 mining for fp32 inputs."""
    L.require_cuda(E_all)
    if E_all.dim() != 2:
        raise ValueError("E_all must be (B_all, D)")
    E_all = E_all.contiguous()
    B_all, D = E_all.shape
    B = B_all - row_offset if num_local is None else int(num_local)
    dev = E_all.device
    lab = _labels_i32(labels_all, dev)
    if lab.numel() != B_all:
        raise ValueError("labels_all has %d entries for %d rows" % (lab.numel(), B_all))
    dt = L.dtype_code(E_all)
    flags = _bh_flags(normalize, detach_gathered, bool(mine_bf16))
    need = L.lib.efm_bh_workspace_bytes(B, B_all, D, dt, flags)
    ws = workspace if workspace is not None and workspace.numel() >= need else L.workspace(need, dev)
    loss = torch.empty(B, dtype=torch.float32, device=dev)
    idx_p = torch.empty(B, dtype=torch.int32, device=dev)
    idx_n = torch.empty(B, dtype=torch.int32, device=dev)
    dE = torch.empty((B_all, D), dtype=torch.float32, device=dev) if want_grad else None
    hg = None if head_grad is None else head_grad.to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        L.check(L.lib.efm_bh_triplet_fwd_bwd(L.ptr(E_all), L.ptr(lab), B_all, int(row_offset), B, D, dt, float(margin),
                                             L.ptr(hg), flags, L.ptr(loss), L.ptr(idx_p), L.ptr(idx_n), L.ptr(dE),
                                             L.ptr(ws), ws.numel(), L.stream()))
    return loss, idx_p, idx_n, dE


class BatchHardStep:
    """The fused batch-hard step for a FIXED shape with everything allocated once (outputs, workspace): `step()` is one
    library call and a handful of Python operations -- what a training loop that runs the same (B, D) every iteration
    wants (the functional `batch_hard_triplet` spends tens of microseconds per call on allocations and checks, about as
    long as the 63 us GPU step itself).  The returned tensors are REUSED by the next call of the same object."""

    def __init__(self, B_all, D, dtype, margin, *, normalize=True, row_offset=0, num_local=None, want_grad=True,
                 detach_gathered=False, mine_bf16=None, device=None):
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.B_all, self.D, self.dtype, self.device = int(B_all), int(D), dtype, dev
        self.B = self.B_all - int(row_offset) if num_local is None else int(num_local)
        self.row_offset, self.margin = int(row_offset), float(margin)
        self.dt = L._DTYPES[dtype]
        self.flags = _bh_flags(normalize, detach_gathered, bool(mine_bf16))
        self.ws = L.workspace(L.lib.efm_bh_workspace_bytes(self.B, self.B_all, self.D, self.dt, self.flags), dev)
        self.loss = torch.empty(self.B, dtype=torch.float32, device=dev)
        self.idx_p = torch.empty(self.B, dtype=torch.int32, device=dev)
        self.idx_n = torch.empty(self.B, dtype=torch.int32, device=dev)
        self.dE = torch.empty((self.B_all, self.D), dtype=torch.float32, device=dev) if want_grad else None
        self._out = (L.ptr(self.loss), L.ptr(self.idx_p), L.ptr(self.idx_n), L.ptr(self.dE), L.ptr(self.ws), self.ws.numel())

    def step(self, E_all, labels_all, head_grad=None):
        """E_all: (B_all, D) contiguous CUDA tensor of the construction dtype; labels_all: (B_all,) int32 CUDA tensor.
        Enqueues the step on the current stream of E_all's device (which must be the current device) and returns
        (loss, idx_p, idx_n, dE) -- the object's own buffers."""
        if E_all.dtype != self.dtype or E_all.shape != (self.B_all, self.D) or not E_all.is_contiguous() or not E_all.is_cuda:
            raise ValueError("E_all must be a contiguous (%d, %d) %s CUDA tensor" % (self.B_all, self.D, self.dtype))
        if labels_all.dtype != torch.int32 or labels_all.numel() != self.B_all or not labels_all.is_contiguous():
            raise ValueError("labels_all must be a contiguous (%d,) int32 tensor" % self.B_all)
        o = self._out
        L.check(L.lib.efm_bh_triplet_fwd_bwd(E_all.data_ptr(), labels_all.data_ptr(), self.B_all, self.row_offset, self.B, self.D,
                                             self.dt, self.margin, L.ptr(head_grad), self.flags, o[0], o[1], o[2], o[3], o[4], o[5],
                                             torch.cuda.current_stream(self.device).cuda_stream))
        return self.loss, self.idx_p, self.idx_n, self.dE


def batch_hard_backward(E_all, idx_p, idx_n, margin, head_grad, *, normalize=True, row_offset=0, detach_gathered=False):
    """Gradient for given mined indices and per-anchor head gradient -> dE_all (B_all, D) fp32."""
    L.require_cuda(E_all)
    E_all = E_all.contiguous()
    B_all, D = E_all.shape
    B = idx_p.numel()
    dev = E_all.device
    flags = _bh_flags(normalize, detach_gathered, False)
    ws = L.workspace(L.lib.efm_bh_backward_workspace_bytes(B_all), dev)
    dE = torch.empty((B_all, D), dtype=torch.float32, device=dev)
    hg = None if head_grad is None else head_grad.to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        L.check(L.lib.efm_bh_backward(L.ptr(E_all), B_all, int(row_offset), B, D, L.dtype_code(E_all), float(margin),
                                      L.ptr(hg), flags, L.ptr(idx_p.contiguous()), L.ptr(idx_n.contiguous()), L.ptr(dE),
                                      L.ptr(ws), ws.numel(), L.stream()))
    return dE


def batch_hard_grad_rows(E_all, idx_p_all, idx_n_all, wgt_all, *, row_offset=0, num_local=None, normalize=True):
    """Exact gradient for rows [row_offset, row_offset+num_local) of the gathered batch from the gathered mining
    results of EVERY anchor (anchor j = row j): wgt_all[j] = 2 * head_grad_j where loss_j > 0, else 0.  This is the
    multi-GPU backward without a dense reduce-scatter (see distributed.global_batch_hard).  Returns (num_local, D) fp32."""
    L.require_cuda(E_all, idx_p_all, idx_n_all, wgt_all)
    E_all = E_all.contiguous()
    B_all, D = E_all.shape
    B = B_all - row_offset if num_local is None else int(num_local)
    dev = E_all.device
    ip = idx_p_all.to(torch.int32).contiguous(); inn = idx_n_all.to(torch.int32).contiguous()
    w = wgt_all.to(torch.float32).contiguous()
    if not (ip.numel() == inn.numel() == w.numel() == B_all):
        raise ValueError("idx_p_all / idx_n_all / wgt_all must have one entry per row of E_all")
    ws = L.workspace(L.lib.efm_bh_grad_rows_workspace_bytes(B_all, B), dev)
    dE = torch.empty((B, D), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib.efm_bh_grad_rows(L.ptr(E_all), B_all, D, L.dtype_code(E_all), _bh_flags(normalize, False, False),
                                       L.ptr(ip), L.ptr(inn), L.ptr(w), int(row_offset), B, L.ptr(dE), L.ptr(ws),
                                       ws.numel(), L.stream()))
    return dE


class _BatchHard(torch.autograd.Function):
    """reduction 'sum'/'mean': gradient is produced by the SAME fused call as the loss (head-grad ones) and
    scaled by the incoming scalar in backward.  reduction 'none': forward only; backward re-runs the cheap
    finalize step with the per-anchor head gradient on the saved indices (no second mining pass)."""

    @staticmethod
    def forward(ctx, E, labels, margin, normalize, reduction, mine_bf16):
        fused = reduction in ("sum", "mean") and E.requires_grad
        loss, idx_p, idx_n, dE = batch_hard_triplet(E, labels, margin, normalize=normalize, want_grad=fused,
                                                    mine_bf16=mine_bf16)
        ctx.mark_non_differentiable(idx_p, idx_n)
        ctx.reduction, ctx.margin, ctx.normalize = reduction, float(margin), normalize
        ctx.in_dtype = E.dtype
        if reduction == "none":
            ctx.save_for_backward(E, idx_p, idx_n)
            return loss, idx_p, idx_n
        ctx.scale = 1.0 / max(1, loss.numel()) if reduction == "mean" else 1.0
        ctx.save_for_backward(dE if fused else E, idx_p, idx_n)
        ctx.fused = fused
        total = loss.sum()
        return (total * ctx.scale if reduction == "mean" else total), idx_p, idx_n

    @staticmethod
    def backward(ctx, g, _gp, _gn):
        t0, idx_p, idx_n = ctx.saved_tensors
        if ctx.reduction == "none":
            dE = batch_hard_backward(t0, idx_p, idx_n, ctx.margin, g, normalize=ctx.normalize)
        else:
            dE = t0 * (g * ctx.scale)
        return dE.to(ctx.in_dtype), None, None, None, None, None


class BatchHardTripletLoss(nn.Module):
    """loss = BatchHardTripletLoss(margin, normalize=True, reduction='none')(embeddings, labels)

    Returns the per-anchor loss vector (reduction 'none', like gluon's TripletLoss) or its sum / mean; the
    mined indices of the last call are kept in `.idx_p` / `.idx_n` (int32, -1 = no candidate)."""

    def __init__(self, margin=0.2, normalize=True, reduction="none", mine_bf16=None):
        super().__init__()
        if reduction not in ("none", "sum", "mean"):
            raise ValueError("reduction must be none|sum|mean")
        self.margin, self.normalize, self.reduction, self.mine_bf16 = margin, normalize, reduction, mine_bf16
        self.idx_p = self.idx_n = None

    def forward(self, embeddings, labels):
        loss, self.idx_p, self.idx_n = _BatchHard.apply(embeddings, labels, self.margin, self.normalize, self.reduction,
                                                        self.mine_bf16)
        return loss
