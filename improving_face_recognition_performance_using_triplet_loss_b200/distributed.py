"""Multi-GPU forms of the two paths that have a real exchange step.

* global batch-hard mining, fused (`PeerExchange`): one process per GPU, CUDA-IPC exchange regions, and ONE step per rank
  with no NCCL call -- extra warps of the Gram + mining kernel copy the peers' rows over NVLink while the tensor pipeline
  mines the blocks that have already arrived; the exact gradient fetches 12 bytes per anchor from the owners instead of
  reduce-scattering a dense dE (csrc/bh.cu, efm_bh_dist_fwd_bwd).
* global batch-hard mining, NCCL form (`global_batch_hard`): all-gather of embeddings + labels (gloo in the CPU tests of the
  host logic), every rank mines its B/P x B strip against the global batch; row extrema are complete per rank, so mining
  needs no reduction.  Exact mode all-gathers 12 bytes per anchor and every rank evaluates the terms that land on its own
  rows; detached mode (reference-faithful, negatives are constants) needs no backward collective.
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


class PeerExchange:
    """Exchange regions of the fused multi-GPU batch-hard step (one process per GPU on ONE node).

    Every rank allocates a region through libefmtl (`efm_xchg_alloc`), the 64-byte CUDA IPC handles travel through
    `torch.distributed` (any backend: they are host bytes), and every rank maps every peer's region.  After that
    `step()` needs no collective at all: the Gram + mining kernel copies the peers' rows over NVLink while it runs
    (csrc/bh.cu, efm_bh_dist_fwd_bwd).  All ranks must call `step()` the same number of times with the same shapes
    and flags."""

    def __init__(self, B_local: int, D: int, dtype: torch.dtype, device=None, group=None):
        import ctypes as C
        import socket

        from . import _lib as L
        self._L, self._C = L, C
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.B, self.D, self.dtype = int(B_local), int(D), dtype
        self.dt = L._DTYPES[dtype]
        nbytes = L.lib.efm_bh_xchg_bytes(self.B, self.world, self.D, self.dt)
        if nbytes == 0 or self.B % 256 or self.D % 8 or self.D > 512 or self.world > 8:
            raise ValueError("fused exchange needs B_local % 256 == 0, D % 8 == 0, D <= 512, world <= 8 "
                             "(got B_local=%d D=%d world=%d)" % (self.B, self.D, self.world))
        self.nbytes = nbytes
        self._own = C.c_void_p()
        self._mapped = []
        self.regions = (C.c_void_p * self.world)()
        with torch.cuda.device(self.device):
            L.check(L.lib.efm_xchg_alloc(nbytes, C.byref(self._own)))
            handle = C.create_string_buffer(64)
            L.check(L.lib.efm_ipc_export(self._own, handle))
            mine = (socket.gethostname(), handle.raw)
            everyone = [mine]
            if self.world > 1:
                everyone = [None] * self.world
                dist.all_gather_object(everyone, mine, group=group)
                if any(h != mine[0] for h, _ in everyone):
                    raise RuntimeError("PeerExchange: all ranks must run on one node (CUDA IPC over NVLink)")
            for r, (_, raw) in enumerate(everyone):
                if r == self.rank:
                    self.regions[r] = self._own.value
                else:
                    p = C.c_void_p()
                    L.check(L.lib.efm_ipc_import(raw, C.byref(p)))
                    self._mapped.append(p)
                    self.regions[r] = p.value
        self.mode = "ipc-pull" if self.world > 1 else "single"
        self.ws = L.workspace(L.lib.efm_bh_dist_workspace_bytes(self.B, self.world), self.device)
        self.steps = 0
        if self.world > 1:
            dist.barrier(group=group)          # nobody reads a region before it exists, is zeroed and is mapped everywhere

    def step(self, E_local, labels_local, margin, *, normalize=True, exact_grad=True, want_grad=True, head_grad=None,
             out=None):
        """One fused step.  Returns (loss (B,), idx_p, idx_n (global rows), dE_local (B, D) fp32 or None).
        `out` = optional preallocated (loss, idx_p, idx_n, dE_local)."""
        L = self._L
        L.require_cuda(E_local)
        if tuple(E_local.shape) != (self.B, self.D) or E_local.dtype != self.dtype:
            raise ValueError("E_local must be (%d, %d) %s" % (self.B, self.D, self.dtype))
        E_local = E_local.contiguous()
        lab = labels_local.reshape(-1)
        if lab.dtype != torch.int32:
            lab = lab.to(torch.int32)
        lab = lab.contiguous()
        dev = E_local.device
        if out is None:
            loss = torch.empty(self.B, dtype=torch.float32, device=dev)
            idx_p = torch.empty(self.B, dtype=torch.int32, device=dev)
            idx_n = torch.empty(self.B, dtype=torch.int32, device=dev)
            dE = torch.empty((self.B, self.D), dtype=torch.float32, device=dev) if want_grad else None
        else:
            loss, idx_p, idx_n, dE = out
        hg = None if head_grad is None else head_grad.to(torch.float32).contiguous()
        flags = (L.BH_NORMALIZE if normalize else 0) | (0 if exact_grad else L.BH_DETACH_GATHERED)
        self.steps += 1
        with torch.cuda.device(dev):
            L.check(L.lib.efm_bh_dist_fwd_bwd(L.ptr(E_local), L.ptr(lab), self.B, self.D, self.dt, float(margin), L.ptr(hg),
                                              flags, self.world, self.rank, self.regions, self.steps, L.ptr(loss),
                                              L.ptr(idx_p), L.ptr(idx_n), L.ptr(dE), L.ptr(self.ws), self.ws.numel(),
                                              L.stream()))
        return loss, idx_p, idx_n, dE

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        """Collective: every rank must call it (peers may still be reading this rank's region otherwise)."""
        if self._own is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        L = self._L
        with torch.cuda.device(self.device):
            for p in self._mapped:
                L.lib.efm_ipc_release(p)
            L.lib.efm_xchg_free(self._own)
        self._mapped, self._own = [], None


def global_batch_hard_fused(exchange: PeerExchange, E_local, labels_local, margin, *, normalize=True, exact_grad=True,
                            head_grad=None):
    """`global_batch_hard` without any NCCL call: the exchange runs inside the Gram kernel over NVLink peer memory."""
    return exchange.step(E_local, labels_local, margin, normalize=normalize, exact_grad=exact_grad, head_grad=head_grad)


def sharded_topk(local_scores: torch.Tensor, local_idx: torch.Tensor, group=None, _merge=None):
    """All-gather per-shard (nq, K) lists (indices already global) and merge -> (nq, K) on every rank."""
    from .gallery import merge_topk
    merge = _merge or merge_topk
    world = dist.get_world_size(group)
    nq, K = local_scores.shape
    ps = all_gather_rows(local_scores, group).reshape(world, nq, K)     # rank-major concatenation along dim 0
    pi = all_gather_rows(local_idx, group).reshape(world, nq, K)
    return merge(ps, pi)
