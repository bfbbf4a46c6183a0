"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the all-gather / reduce-scatter plumbing of global
batch-hard mining and the sharded-gallery merge.  The device operator is replaced by the oracle through the
`_impl` / `_merge` hooks, so what is tested here is exactly the code that runs around the CUDA calls on N GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _oracle_impl(E_all, y_all, margin, *, normalize=True, row_offset=0, num_local=None, want_grad=True,
                 detach_gathered=False, mine_bf16=None, **_):
    import oracle as O
    E = E_all.double().numpy(); y = y_all.numpy()
    B = num_local
    r = O.batch_hard_triplet(E[row_offset:row_offset + B], y[row_offset:row_offset + B], margin, normalize=normalize,
                             E_all=E, labels_all=y, row_offset=row_offset)
    dE = r["dE_all"].copy()
    if detach_gathered:                                              # only local rows keep their gradient
        r2 = dE.copy(); r2[:] = 0
        # recompute with role masking: anchors always local; positives/negatives only if local
        X = r["X_all"]; w = (r["loss"] > 0).astype(np.float64)
        dX = np.zeros_like(X)
        for i in range(B):
            if w[i] == 0:
                continue
            gi, p, n = row_offset + i, r["idx_p"][i], r["idx_n"][i]
            dX[gi] += 2 * (X[n] - X[p])
            if row_offset <= p < row_offset + B:
                dX[p] += 2 * (X[p] - X[gi])
            if row_offset <= n < row_offset + B:
                dX[n] += 2 * (X[gi] - X[n])
        nrm = np.sqrt((E * E).sum(1, keepdims=True)) if normalize else 1.0
        dE = (dX - X * (X * dX).sum(1, keepdims=True)) / nrm if normalize else dX
    return (torch.from_numpy(r["loss"]).float(), torch.from_numpy(r["idx_p"]).int(), torch.from_numpy(r["idx_n"]).int(),
            torch.from_numpy(dE).float())


def _oracle_grad_rows(E_all, ip_all, in_all, w_all, *, row_offset=0, num_local=None, normalize=True):
    """Row-sliced exact gradient (oracle): every anchor's contribution that lands on rows [row_offset, +num_local)."""
    E = E_all.double().numpy(); ip = ip_all.numpy(); inn = in_all.numpy(); w = w_all.double().numpy()
    nrm = np.sqrt((E * E).sum(1, keepdims=True)) if normalize else np.ones((E.shape[0], 1))
    X = E / nrm
    dX = np.zeros_like(X)
    for j in range(E.shape[0]):
        if w[j] == 0:
            continue
        p, n = ip[j], inn[j]
        dX[j] += w[j] * (X[n] - X[p]); dX[p] += w[j] * (X[p] - X[j]); dX[n] += w[j] * (X[j] - X[n])
    dE = (dX - X * (X * dX).sum(1, keepdims=True)) / nrm if normalize else dX
    return torch.from_numpy(dE[row_offset:row_offset + num_local]).float()


def _numpy_merge(ps, pi):
    parts, nq, K = ps.shape
    out_s = torch.zeros(nq, K); out_i = torch.full((nq, K), -1, dtype=torch.int32)
    for q in range(nq):
        c = [(float(ps[p, q, k]), int(pi[p, q, k])) for p in range(parts) for k in range(K) if int(pi[p, q, k]) >= 0]
        c.sort(key=lambda t: (-t[0], t[1]))
        for k, (s, i) in enumerate(c[:K]):
            out_s[q, k] = s; out_i[q, k] = i
    return out_s, out_i


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from improving_face_recognition_performance_using_triplet_loss_b200 import distributed as D
    from util import clustered
    E, y = clustered(32, 8, seed=1)
    B = 32 // world
    El = torch.from_numpy(E[rank * B:(rank + 1) * B]); yl = torch.from_numpy(y[rank * B:(rank + 1) * B])
    res = {}
    for exact in (True, False):
        loss, ip, inn, dE = D.global_batch_hard(El, yl, 1.9, normalize=True, exact_grad=exact, _impl=_oracle_impl,
                                                _grad_rows=_oracle_grad_rows)
        res["exact" if exact else "detached"] = (loss.numpy(), ip.numpy(), inn.numpy(), dE.numpy())
    # sharded gallery: every rank scores its shard, lists carry global indices
    rng = np.random.default_rng(0)
    G = rng.standard_normal((40, 8)).astype(np.float32); Q = G[[3, 25]] + 0.01
    lo, hi = D.shard_bounds(40, world, rank)
    sims = (Q @ G[lo:hi].T)
    K = 3
    top = np.argsort(-sims, axis=1)[:, :K]
    ls = torch.from_numpy(np.take_along_axis(sims, top, 1)); li = torch.from_numpy((top + lo).astype(np.int32))
    ms, mi = D.sharded_topk(ls, li, _merge=_numpy_merge)
    res["topk"] = (ms.numpy(), mi.numpy(), np.argsort(-(Q @ G.T), axis=1)[:, :K])
    out[rank] = res
    dist.destroy_process_group()


def test_global_batch_hard_and_sharded_topk_world2():
    import oracle as O
    from util import clustered
    world = 2
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    E, y = clustered(32, 8, seed=1)
    full = O.batch_hard_triplet(E, y, 1.9, normalize=True)
    assert (full["loss"] > 0).sum() > 8
    loss = np.concatenate([out[r]["exact"][0] for r in range(world)])
    ip = np.concatenate([out[r]["exact"][1] for r in range(world)])
    dE = np.concatenate([out[r]["exact"][3] for r in range(world)])
    assert np.allclose(loss, full["loss"], atol=1e-5) and np.array_equal(ip, full["idx_p"])
    assert np.allclose(dE, full["dE_all"], atol=1e-5)                   # sparse exchange == single-process gradient
    # detached mode: each rank's gradient covers only its own rows and ignores remote positive/negative roles
    for r in range(world):
        assert out[r]["detached"][3].shape == (16, 8)
    for r in range(world):
        ms, mi, ref = out[r]["topk"]
        assert np.array_equal(mi, ref)


def test_shard_bounds_cover_everything():
    from improving_face_recognition_performance_using_triplet_loss_b200.distributed import shard_bounds
    for n, w in ((1000000, 8), (13233, 4), (7, 8)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
