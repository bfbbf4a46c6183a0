"""Cosine verification scoring and its statistics.

`cosine_dist(anc, pos, neg, batch_size)` keeps the reference signature (train_efm.py:26-34) but does the 2B
dot/norm/div chains in one kernel launch each instead of ~12B tiny MXNet ops + 2B host syncs.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


def cosine_pairs(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """dot(a_i, b_i) / (|a_i| * |b_i|) per row -> (P,) fp32."""
    L.require_cuda(a, b)
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("cosine_pairs: shape/dtype mismatch %s %s" % (a.shape, b.shape))
    P = a.shape[0]
    a2, b2 = a.reshape(P, -1).contiguous(), b.reshape(P, -1).contiguous()
    out = torch.empty(P, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        L.check(L.lib.efm_cosine_pairs(L.ptr(a2), L.ptr(b2), P, a2.shape[1], L.dtype_code(a2), L.ptr(out), L.stream()))
    return out


def cosine_pairs_indexed(E: torch.Tensor, ia: torch.Tensor, ib: torch.Tensor) -> torch.Tensor:
    """LFW-style pair list: cos(E[ia[k]], E[ib[k]]) -> (P,) fp32."""
    L.require_cuda(E, ia, ib)
    E2 = E.reshape(E.shape[0], -1).contiguous()
    ia = ia.to(torch.int32).contiguous(); ib = ib.to(torch.int32).contiguous()
    P = ia.numel()
    if P and (int(ia.max()) >= E2.shape[0] or int(ib.max()) >= E2.shape[0] or int(ia.min()) < 0 or int(ib.min()) < 0):
        raise ValueError("cosine_pairs_indexed: pair index out of range")
    out = torch.empty(P, dtype=torch.float32, device=E.device)
    with torch.cuda.device(E.device):
        L.check(L.lib.efm_cosine_pairs_indexed(L.ptr(E2), L.ptr(ia), L.ptr(ib), P, E2.shape[1], L.dtype_code(E2),
                                               L.ptr(out), L.stream()))
    return out


def cosine_dist(anc, pos, neg, batch_size, as_lists=True):
    """Reference signature.  as_lists=True returns two Python lists of shape-(1,) tensors exactly like the
    reference (so `pos_dist[v].item()` call sites keep working); as_lists=False returns two (B,) tensors."""
    p = cosine_pairs(anc[:batch_size], pos[:batch_size])
    n = cosine_pairs(anc[:batch_size], neg[:batch_size])
    if not as_lists:
        return p, n
    return list(p.reshape(-1, 1).unbind(0)), list(n.reshape(-1, 1).unbind(0))


def cosine_allpairs(A: torch.Tensor, B: torch.Tensor | None = None) -> torch.Tensor:
    """(M, N) fp32 matrix of cos(A_i, B_j) on the tcgen05 path (operands rounded to bf16, fp32 accumulate)."""
    B = A if B is None else B
    L.require_cuda(A, B)
    if A.dtype != B.dtype or A.shape[1] != B.shape[1]:
        raise ValueError("cosine_allpairs: dtype/width mismatch")
    A2, B2 = A.contiguous(), B.contiguous()
    M, D = A2.shape
    N = B2.shape[0]
    ld = (N + 31) // 32 * 32                                # 128-byte aligned rows: every store hits whole sectors
    buf = torch.empty((M, ld), dtype=torch.float32, device=A.device)
    dt = L.dtype_code(A2)
    ws = L.workspace(L.lib.efm_cosine_allpairs_workspace_bytes(M, N, D, dt), A.device)
    with torch.cuda.device(A.device):
        L.check(L.lib.efm_cosine_allpairs(L.ptr(A2), L.ptr(B2), M, N, D, dt, L.ptr(buf), ld, L.ptr(ws), ws.numel(), L.stream()))
    return buf[:, :N]                                       # view into the padded buffer (call .contiguous() if needed)


def write_cosine_csv(path, pos, neg, mode="a+"):
    """cosine_similarity.csv rows "pos neg" exactly as train_efm.py:250-255 writes them: csv.writer(delimiter=' ') over
    `asscalar()` values, which are numpy.float32 scalars -- csv stringifies those with str(), i.e. the SHORTEST float32
    repr ('0.70710677', not the 17-digit repr of the widened double), '\r\n' line ends.  One host copy per call."""
    import csv
    p = pos.detach().float().cpu().numpy().reshape(-1) if isinstance(pos, torch.Tensor) else np.asarray(pos, dtype=np.float32).reshape(-1)
    n = neg.detach().float().cpu().numpy().reshape(-1) if isinstance(neg, torch.Tensor) else np.asarray(neg, dtype=np.float32).reshape(-1)
    with open(path, mode, newline="") as f:
        w = csv.writer(f, delimiter=" ")
        for a, b in zip(p, n):
            w.writerow([np.float32(a), np.float32(b)])


# ---- draw_cos_dis_real.py statistics (plotting itself is out of scope) -----------------------------------
def read_csv(path="cosine_similarity.csv", desire_epoch=1):
    """draw_cos_dis_real.py:9-21."""
    pos, neg = [], []
    with open(path, "r") as f:
        data = f.readlines()
    for i, line in enumerate(data):
        if i > (len(data) - int(len(data) / desire_epoch)):
            parts = line.split(" ")
            pos.append(float(parts[0])); neg.append(float(parts[1]))
    return pos, neg


def calc_counts(dis_list, bins=100):
    """draw_cos_dis_real.py:4-7: np.histogram(x, bins=100) -> (count, bin_edges); the counting runs on the
    GPU with numpy's exact uniform-bin rule, edges are numpy's linspace over [min, max]."""
    x = dis_list if isinstance(dis_list, torch.Tensor) else torch.tensor(np.asarray(dis_list, dtype=np.float32))
    if not x.is_cuda:
        x = x.cuda()
    x = x.detach().reshape(-1).to(torch.float32).contiguous()
    if x.numel() == 0:
        raise ValueError("calc_counts: empty input")
    lo, hi = float(x.min()), float(x.max())
    if lo == hi:                                            # numpy widens a degenerate range by +-0.5
        lo, hi = lo - 0.5, hi + 0.5
    edges = np.linspace(lo, hi, bins + 1, endpoint=True)
    d_edges = torch.from_numpy(edges).to(x.device)
    counts = torch.zeros(bins, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.lib.efm_histogram(L.ptr(x), x.numel(), L.ptr(d_edges), bins, L.ptr(counts), L.stream()))
    return counts.cpu().numpy(), edges


def calc_lines(pos_dis, neg_dis):
    """draw_cos_dis_real.py:23-34: PDFs, positive CDF, negative 1-CDF, and the bin edges."""
    pc, pb = calc_counts(pos_dis); nc, nb = calc_counts(neg_dis)
    pl = pc / sum(pc); nl = nc / sum(nc)
    return pl, nl, np.cumsum(pl), 1 - np.cumsum(nl), pb, nb
