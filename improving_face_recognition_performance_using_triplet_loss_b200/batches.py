"""Triplet batch construction: the caller side of the loss head.

* `define_pos` + `TripletPairIter` reproduce the batch layout contract of the reference's loader (train_efm.py:37-114,
  the same in pre-trained_efm_v3.py:34-111): a pre-pass fixes, per identity, the FIRST sample seen as that identity's
  positive; every batch then carries 2B rows -- rows 0..B-1 the anchors, rows B..2B-1 the anchors' fixed positives in the
  same order, labels duplicated -- so that `anc = fc[0:B]`, `pos = fc[B:2B]` (train_efm.py:232-233).  The reference
  builds every batch with 4B host round trips (`asnumpy` / `asscalar` per row); here the two index vectors are built
  once on the host and a batch is one gather on the device that holds the data (`torch.index_select`).
* `pk_batches`: P identities x K samples per batch, the sampler online batch-hard mining wants (north-star
  extension; no reference code).

Pure data plumbing (indexing only): works on whatever device the tensors live on.  Pinned against the reference's own
`define_pos` / `DataIter` run in the build container (tests/golden/ref_python.npz, tests/test_batches_cpu.py).
"""
from __future__ import annotations

import numpy as np
import torch


def define_pos(labels) -> dict:
    """train_efm.py:37-45: identity -> index of the first sample with that identity, in iteration order (dict order =
    order of first appearance).  Labels are compared as int(float label) like the reference."""
    lab = np.asarray(torch.as_tensor(labels).detach().cpu().numpy() if isinstance(labels, torch.Tensor) else labels).reshape(-1)
    pos = {}
    for i, v in enumerate(lab.tolist()):
        k = int(v)
        if k not in pos:
            pos[k] = i
    return pos


class TripletPairIter:
    """DataIter(data_iter, length, pos_img, batch_size, dshape) of train_efm.py:62-114 over an in-memory dataset.

    data: (N, ...) tensor, labels: (N,) tensor; `length` batches of `batch_size` anchors are emitted in dataset order
    (the reference's `start += batch_size`).  Each item is (data_2B, label_2B): anchors then their positives, labels twice
    (float32, as MXNet iterators carry them)."""

    def __init__(self, data: torch.Tensor, labels, batch_size: int, length: int | None = None, pos: dict | None = None):
        self.data = data
        self.labels = torch.as_tensor(labels).reshape(-1)
        n = self.labels.numel()
        if data.shape[0] != n:
            raise ValueError("data has %d rows, labels %d" % (data.shape[0], n))
        self.batch_size = int(batch_size)
        self.length = n // self.batch_size if length is None else int(length)
        if self.length * self.batch_size > n:
            raise ValueError("%d batches of %d need more than the %d samples given" % (self.length, self.batch_size, n))
        pos = define_pos(self.labels) if pos is None else pos
        lab_host = self.labels.detach().cpu().numpy().reshape(-1)
        try:
            pidx = np.array([pos[int(v)] for v in lab_host.tolist()], dtype=np.int64)
        except KeyError as e:
            raise KeyError("identity %s has no positive sample (define_pos was built on another dataset)" % e) from None
        dev = data.device
        self._anchor_idx = torch.arange(n, dtype=torch.long, device=dev)
        self._pos_idx = torch.from_numpy(pidx).to(dev)
        self._labels_f32 = self.labels.to(torch.float32).to(dev)

    def __len__(self):
        return self.length

    def batch_indices(self, i: int):
        """Row indices of batch i into the dataset: (2B,) = [anchors; positives]."""
        s = slice(i * self.batch_size, (i + 1) * self.batch_size)
        return torch.cat([self._anchor_idx[s], self._pos_idx[s]])

    def __iter__(self):
        for i in range(self.length):
            idx = self.batch_indices(i)
            lab = self._labels_f32[i * self.batch_size:(i + 1) * self.batch_size]
            yield self.data.index_select(0, idx), torch.cat([lab, lab])

    def reset(self):
        """The reference resets its underlying iterator (train_efm.py:112-114); nothing to do for an in-memory dataset."""


def pk_batches(labels, P: int, K: int, n_batches: int, seed: int = 0):
    """P x K sampler for batch-hard mining: every batch holds K samples of each of P distinct identities (sampled without
    replacement among identities that own at least K samples; samples of an identity without replacement).  Returns a
    (n_batches, P*K) int64 tensor of dataset row indices, rows grouped identity by identity."""
    lab = np.asarray(torch.as_tensor(labels).detach().cpu().numpy()).reshape(-1)
    ids, inv = np.unique(lab, return_inverse=True)
    members = [np.flatnonzero(inv == k) for k in range(len(ids))]
    eligible = np.array([k for k, m in enumerate(members) if len(m) >= K])
    if len(eligible) < P:
        raise ValueError("only %d identities own >= %d samples, need %d per batch" % (len(eligible), K, P))
    rng = np.random.default_rng(seed)
    out = np.empty((n_batches, P * K), dtype=np.int64)
    for b in range(n_batches):
        chosen = rng.choice(eligible, size=P, replace=False)
        out[b] = np.concatenate([rng.choice(members[k], size=K, replace=False) for k in chosen])
    return torch.from_numpy(out)
