"""Multi-GPU forms of the two paths that have a real exchange step.

* global batch-hard mining: all-gather of embeddings + labels (NCCL over NVLink on GPUs, gloo in CPU tests of
  the host logic), every rank mines its B/P x B strip against the global batch; row extrema are complete per
  rank, so mining needs no reduction.  Exact mode reduce-scatters dE_all so the result equals the single-GPU
  loss on the concatenated batch WITHOUT a dense reduce-scatter (12 bytes per anchor are all-gathered instead and
  every rank evaluates the terms that land on its own rows); detached mode (reference-faithful, negatives are
  constants) needs no backward collective.
* sharded gallery top-k: each rank owns a contiguous row shard, local top-k with global indices, all-gather of
  (score, idx) lists and a K-way merge on every rank.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of n rows for `rank` (first n % world shards one row longer)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate equally-sized row blocks of all ranks (rank order)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def global_batch_hard(E_local: torch.Tensor, labels_local: torch.Tensor, margin: float, *, normalize=True,
                      exact_grad=True, mine_bf16=None, head_grad=None, group=None, _impl=None, _grad_rows=None):
    """Batch-hard triplet loss over the global batch.  Returns (loss_local (B,), idx_p, idx_n (global row
    indices), dE_local (B, D) fp32).

    exact_grad=True  -> the gradient equals the single-GPU gradient of the concatenated batch.  No dense
        reduce-scatter: every rank all-gathers 12 bytes per anchor (idx_p, idx_n, weight) and evaluates the
        positive / negative-role terms that land on ITS rows from the already gathered embeddings.
    exact_grad=False -> gathered rows are constants (mirrors the reference's detached negatives): no backward collective.
    `_impl` / `_grad_rows` let the CPU tests substitute the single-process operators."""
    from .loss import batch_hard_grad_rows, batch_hard_triplet
    impl = _impl or batch_hard_triplet
    grad_rows = _grad_rows or batch_hard_grad_rows
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = E_local.shape[0]
    E_all = all_gather_rows(E_local, group)
    y_all = all_gather_rows(labels_local.reshape(-1).to(torch.int32), group)
    if not exact_grad or world == 1:
        loss, idx_p, idx_n, dE_all = impl(E_all, y_all, margin, normalize=normalize, row_offset=rank * B, num_local=B,
                                          want_grad=True, detach_gathered=not exact_grad, mine_bf16=mine_bf16,
                                          head_grad=head_grad)
        return loss, idx_p, idx_n, dE_all[rank * B:(rank + 1) * B]
    loss, idx_p, idx_n, _ = impl(E_all, y_all, margin, normalize=normalize, row_offset=rank * B, num_local=B,
                                 want_grad=False, mine_bf16=mine_bf16)
    w = torch.where(loss > 0, torch.full_like(loss, 2.0), torch.zeros_like(loss))       # relu'(0) = 0
    if head_grad is not None:
        w = w * head_grad.to(w.dtype)
    packed = torch.stack([idx_p.to(torch.int32), idx_n.to(torch.int32), w.to(torch.float32).view(torch.int32)])  # (3, B)
    allp = all_gather_rows(packed.unsqueeze(0), group)                                   # (world, 3, B)
    ip_all = allp[:, 0].reshape(-1); in_all = allp[:, 1].reshape(-1)
    w_all = allp[:, 2].reshape(-1).contiguous().view(torch.float32)
    dE_local = grad_rows(E_all, ip_all, in_all, w_all, row_offset=rank * B, num_local=B, normalize=normalize)
    return loss, idx_p, idx_n, dE_local


def sharded_topk(local_scores: torch.Tensor, local_idx: torch.Tensor, group=None, _merge=None):
    """All-gather per-shard (nq, K) lists (indices already global) and merge -> (nq, K) on every rank."""
    from .gallery import merge_topk
    merge = _merge or merge_topk
    world = dist.get_world_size(group)
    nq, K = local_scores.shape
    ps = all_gather_rows(local_scores, group).reshape(world, nq, K)     # rank-major concatenation along dim 0
    pi = all_gather_rows(local_idx, group).reshape(world, nq, K)
    return merge(ps, pi)
