"""CPU restatement (numpy) of the reference's embedding-metric hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py for who may import this.
Every function cites the reference file:line it follows (paths relative to the
reference repo root).  MXNet itself is not vendored in the reference and is not
installable here; statements about MXNet op semantics are from MXNet 1.x public
source and are marked [MXNet].  For those functions parity is UNPINNED (no golden
vector exists in the reference); the native scoring functions are pinned against
the reference's own C++ built into oracle/_ref (see oracle/Makefile).
"""
from __future__ import annotations

import math
import random as _pyrandom

import numpy as np

__all__ = [
    "bf16_round", "triplet_loss_fwd", "triplet_loss_bwd", "sample_negatives",
    "pairwise_sqdist", "batch_hard_mine", "batch_hard_triplet", "batch_hard_triplet_fast", "batch_hard_strip_fast",
    "norm", "l2norm_fwd", "l2norm_bwd", "frob_fwd", "frob_bwd",
    "mx_maximum", "mx_minimum", "efm3_fwd", "efm3_bwd", "mfm2_fwd", "mfm2_bwd",
    "cosine_pairs", "cosine_dist_literal", "cosine_allpairs", "format_cosine_csv",
    "read_cosine_csv", "calc_counts", "calc_lines", "simd_dot", "simd_dot_rows",
    "gallery_scores", "gallery_argmax", "gallery_topk", "compare_face_person",
    "compare_face_db", "format_feature_rows", "format_label_rows",
    "lfw_distance", "lfw_calculate_accuracy", "lfw_calculate_roc", "kfold_indices",
    "reference_literal_step",
]

F32 = np.float32


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) and return as fp32 (same rule as
    __float2bfloat16_rn / torch .to(bfloat16)).  NaN stays NaN."""
    x = np.ascontiguousarray(x, dtype=F32)
    u = x.view(np.uint32).astype(np.uint64)
    bias = ((u >> 16) & 1) + 0x7FFF
    r = ((u + bias) & 0xFFFF0000).astype(np.uint32)
    out = r.view(F32).copy()
    nan = np.isnan(x)
    if nan.any():
        out[nan] = np.nan
    return out.reshape(x.shape)


# --------------------------------------------------------------------------------------
# a1  gluon.loss.TripletLoss   (call sites pre-trained_efm_v3.py:183,210-212;
#     train_efm.py:210,241; final_efm.py:220,259)
# --------------------------------------------------------------------------------------
def triplet_loss_fwd(a, p, n, margin, dtype=F32):
    """loss_i = relu( sum_d[(p-a)^2 - (n-a)^2] + margin ), shape (B,), no mean.
    [MXNet] gluon.loss.TripletLoss.hybrid_forward: sum(square(pos-pred) -
    square(neg-pred), axis=batch_axis, exclude=True); relu(loss + margin)."""
    a = np.asarray(a, dtype=dtype); p = np.asarray(p, dtype=dtype); n = np.asarray(n, dtype=dtype)
    t = np.square(p - a) - np.square(n - a)
    s = t.reshape(t.shape[0], -1).sum(axis=1, dtype=dtype)
    return np.maximum(s + dtype(margin), dtype(0)).astype(dtype)


def triplet_loss_bwd(a, p, n, margin, head_grad=None, dtype=F32):
    """Gradients of sum_i head_grad_i * loss_i (loss.backward() uses head-grad ones,
    pre-trained_efm_v3.py:211).  relu'(0) = 0 [MXNet relu backward: out > 0].
    Returns (da, dp, dn)."""
    a = np.asarray(a, dtype=dtype); p = np.asarray(p, dtype=dtype); n = np.asarray(n, dtype=dtype)
    loss = triplet_loss_fwd(a, p, n, margin, dtype)
    w = np.ones_like(loss) if head_grad is None else np.asarray(head_grad, dtype=dtype)
    w = (w * (loss > 0)).reshape((-1,) + (1,) * (a.ndim - 1)).astype(dtype)
    two = dtype(2)
    return (two * (n - p) * w).astype(dtype), (two * (p - a) * w).astype(dtype), (two * (a - n) * w).astype(dtype)


# --------------------------------------------------------------------------------------
# a2  random negative selection (train_efm.py:234-239 train pool = anchors only;
#     :268-273 validation pool = all 2B rows; pre-trained_efm_v3.py:202-207,233-238)
# --------------------------------------------------------------------------------------
def sample_negatives(labels, batch_size, pool="train", rng=None):
    """Replays the reference's rejection sampling call-for-call:
        j = random.randint(0, hi); while int(label[j]) == int(label[i]): redraw
    hi = B-1 (train) or 2B-1 (val/test).  `rng` is a random.Random (the reference uses
    the unseeded module-level generator; pass random.Random(seed) to replay)."""
    rng = rng or _pyrandom
    hi = batch_size - 1 if pool == "train" else 2 * batch_size - 1
    lab = [int(float(v)) for v in np.asarray(labels).reshape(-1)]
    if len(set(lab[: hi + 1])) < 2:
        raise ValueError("single-identity pool: the reference loops forever here")
    out = np.empty(batch_size, dtype=np.int64)
    for i in range(batch_size):
        j = rng.randint(0, hi)
        while lab[j] == lab[i]:
            j = rng.randint(0, hi)
        out[i] = j
    return out


# --------------------------------------------------------------------------------------
# a2' batch-hard mining (north_star; NO reference code -> parity unpinned)
# --------------------------------------------------------------------------------------
def pairwise_sqdist(E, E_all, dtype=np.float64):
    """d_ij = sum_d (e_i - e_j)^2 evaluated directly (no Gram cancellation)."""
    E = np.asarray(E, dtype=dtype); E_all = np.asarray(E_all, dtype=dtype)
    out = np.empty((E.shape[0], E_all.shape[0]), dtype=dtype)
    step = max(1, (1 << 24) // max(1, E_all.shape[0] * E.shape[1]))
    for s in range(0, E.shape[0], step):
        diff = E[s:s + step, None, :] - E_all[None, :, :]
        out[s:s + step] = np.einsum("ijk,ijk->ij", diff, diff)
    return out


def batch_hard_mine(E, labels, E_all=None, labels_all=None, row_offset=0, dist=None):
    """For local anchor i (global row row_offset+i):
         p*(i) = argmax_{j != row_offset+i, y_j == y_i} d_ij
         n*(i) = argmin_{y_j != y_i} d_ij
    squared Euclidean d, lowest-index tie-break, -1 when no candidate exists."""
    if E_all is None:
        E_all, labels_all = E, labels
    labels = np.asarray(labels).reshape(-1).astype(np.int64)
    labels_all = np.asarray(labels_all).reshape(-1).astype(np.int64)
    d = pairwise_sqdist(E, E_all) if dist is None else np.asarray(dist)
    B, N = d.shape
    same = labels[:, None] == labels_all[None, :]
    rows = np.arange(B)
    not_self = np.ones((B, N), dtype=bool)
    self_col = rows + row_offset
    ok = self_col < N
    not_self[rows[ok], self_col[ok]] = False
    pos_mask = same & not_self
    neg_mask = ~same
    dp = np.where(pos_mask, d, -np.inf)
    dn = np.where(neg_mask, d, np.inf)
    idx_p = np.argmax(dp, axis=1).astype(np.int64)     # first occurrence == lowest index
    idx_n = np.argmin(dn, axis=1).astype(np.int64)
    idx_p[~pos_mask.any(axis=1)] = -1
    idx_n[~neg_mask.any(axis=1)] = -1
    return idx_p, idx_n


def batch_hard_triplet(E, labels, margin, normalize=True, E_all=None, labels_all=None,
                       row_offset=0, head_grad=None, mine_bf16=False):
    """Full batch-hard head: (optional) per-row L2 normalise -> mine -> a1 on
    (e_i, e_p*, e_n*) -> gradient w.r.t. the (un-normalised) embeddings.

    Single-process form: E_all is E (or the all-gathered global batch whose rows
    [row_offset, row_offset+B) are E).  Returns dict(loss, idx_p, idx_n, dE_all) where
    dE_all has E_all's shape (roles anchor/positive/negative all receive gradient,
    i.e. the exact single-GPU-equivalent form of SURVEY 8e).
    mine_bf16: the contract of the tensor-core path -- the MINING rows are the raw rows rounded to
    bf16 (bf16 inputs: the rows themselves); with normalize the mining metric is the cosine of the
    mining rows, (xm_i . xm_j) / (|e_i| |e_j|) with the norms of the original rows (hardest positive =
    smallest cosine, hardest negative = largest; the same ordering as the squared distance of the
    normalised rows), without it their squared distance.  Loss and gradient use the fp32 rows."""
    E = np.asarray(E, dtype=np.float64)
    if E_all is None:
        E_all, labels_all = E, labels
    E_all = np.asarray(E_all, dtype=np.float64)
    if normalize:
        nrm_all = np.sqrt((E_all * E_all).sum(axis=1, keepdims=True))
        X_all = E_all / nrm_all
    else:
        X_all = E_all
    B = E.shape[0]
    X = X_all[row_offset:row_offset + B]
    if mine_bf16:
        Xm_all = bf16_round(E_all.astype(F32)).astype(np.float64)
        Xm = Xm_all[row_offset:row_offset + B]
        if normalize:
            cos = (Xm @ Xm_all.T) / (nrm_all[row_offset:row_offset + B] * nrm_all.reshape(1, -1))
            idx_p, idx_n = batch_hard_mine(Xm, labels, Xm_all, labels_all, row_offset, dist=-cos)
        else:
            idx_p, idx_n = batch_hard_mine(Xm, labels, Xm_all, labels_all, row_offset)
    else:
        idx_p, idx_n = batch_hard_mine(X, labels, X_all, labels_all, row_offset)
    valid = (idx_p >= 0) & (idx_n >= 0)
    ip = np.where(valid, idx_p, 0); inn = np.where(valid, idx_n, 0)
    P = X_all[ip]; Nn = X_all[inn]
    s = ((P - X) ** 2 - (Nn - X) ** 2).sum(axis=1) + margin
    loss = np.where(valid, np.maximum(s, 0.0), 0.0)
    w = np.ones(B) if head_grad is None else np.asarray(head_grad, dtype=np.float64)
    w = w * (loss > 0)
    dX_all = np.zeros_like(X_all)
    rows = np.arange(B) + row_offset
    np.add.at(dX_all, rows, 2.0 * (Nn - P) * w[:, None])
    np.add.at(dX_all, ip, 2.0 * (P - X) * w[:, None])
    np.add.at(dX_all, inn, 2.0 * (X - Nn) * w[:, None])
    if normalize:
        dE_all = (dX_all - X_all * (X_all * dX_all).sum(axis=1, keepdims=True)) / nrm_all
    else:
        dE_all = dX_all
    return dict(loss=loss, idx_p=idx_p, idx_n=idx_n, dE_all=dE_all, X_all=X_all, s=np.where(valid, s, -np.inf))


def batch_hard_triplet_fast(E, labels, margin, normalize=True):
    """Vectorised fp32 port (BLAS sgemm Gram + argmax/argmin + hinge + scatter) used as the
    timed CPU baseline ("port") -- same arithmetic as batch_hard_triplet in fp32 with the
    Gram identity d = |a|^2 + |b|^2 - 2ab for mining."""
    E = np.ascontiguousarray(E, dtype=F32)
    labels = np.asarray(labels).reshape(-1)
    if normalize:
        nrm = np.sqrt((E * E).sum(axis=1, keepdims=True))
        X = E / nrm
    else:
        X = E
    sq = (X * X).sum(axis=1)
    d = sq[:, None] + sq[None, :] - F32(2) * (X @ X.T)
    same = labels[:, None] == labels[None, :]
    B = E.shape[0]
    dp = np.where(same, d, -np.inf); dp[np.arange(B), np.arange(B)] = -np.inf
    dn = np.where(same, np.inf, d)
    idx_p = dp.argmax(axis=1); idx_n = dn.argmin(axis=1)
    valid = np.isfinite(dp[np.arange(B), idx_p]) & np.isfinite(dn[np.arange(B), idx_n])
    P = X[idx_p]; Nn = X[idx_n]
    s = (np.square(P - X) - np.square(Nn - X)).sum(axis=1) + F32(margin)
    loss = np.where(valid, np.maximum(s, 0), 0).astype(F32)
    w = (loss > 0).astype(F32)[:, None]
    dX = F32(2) * (Nn - P) * w
    np.add.at(dX, idx_p, F32(2) * (P - X) * w)
    np.add.at(dX, idx_n, F32(2) * (X - Nn) * w)
    if normalize:
        dE = (dX - X * (X * dX).sum(axis=1, keepdims=True)) / nrm
    else:
        dE = dX
    return loss, np.where(valid, idx_p, -1), np.where(valid, idx_n, -1), dE


def batch_hard_strip_fast(E_all, labels_all, row_offset, num_local, margin, normalize=True):
    """fp32 port for `num_local` anchors [row_offset, row_offset+num_local) mined against the whole (gathered) batch:
    the per-rank work of global batch-hard mining.  Returns (loss, idx_p, idx_n, dE_all)."""
    E_all = np.ascontiguousarray(E_all, dtype=F32)
    y = np.asarray(labels_all).reshape(-1)
    if normalize:
        nrm = np.sqrt((E_all * E_all).sum(axis=1, keepdims=True))
        X = E_all / nrm
    else:
        X = E_all
    sl = slice(row_offset, row_offset + num_local)
    Xa = X[sl]
    sq = (X * X).sum(axis=1)
    d = sq[sl, None] + sq[None, :] - F32(2) * (Xa @ X.T)
    same = y[sl, None] == y[None, :]
    rows = np.arange(num_local)
    dp = np.where(same, d, -np.inf); dp[rows, rows + row_offset] = -np.inf
    dn = np.where(same, np.inf, d)
    idx_p = dp.argmax(axis=1); idx_n = dn.argmin(axis=1)
    valid = np.isfinite(dp[rows, idx_p]) & np.isfinite(dn[rows, idx_n])
    P = X[idx_p]; Nn = X[idx_n]
    s = (np.square(P - Xa) - np.square(Nn - Xa)).sum(axis=1) + F32(margin)
    loss = np.where(valid, np.maximum(s, 0), 0).astype(F32)
    w = (loss > 0).astype(F32)[:, None]
    dX = np.zeros_like(X)
    dX[sl] += F32(2) * (Nn - P) * w
    np.add.at(dX, idx_p, F32(2) * (P - Xa) * w)
    np.add.at(dX, idx_n, F32(2) * (Xa - Nn) * w)
    dE = (dX - X * (X * dX).sum(axis=1, keepdims=True)) / nrm if normalize else dX
    return loss, np.where(valid, idx_p, -1), np.where(valid, idx_n, -1), dE


# --------------------------------------------------------------------------------------
# a3  normalisation (final_efm.py:240-243, extract_feacture_v2.py:70,97 per row;
#     train_efm.py:241 whole tensor)
# --------------------------------------------------------------------------------------
def norm(x, dtype=F32):
    """[MXNet] mx.nd.norm(x) with no axis: sqrt(sum of squares of ALL elements), shape (1,)."""
    x = np.asarray(x, dtype=dtype)
    return np.sqrt((x * x).sum(dtype=dtype)).reshape(1).astype(dtype)


def l2norm_fwd(x, dtype=F32):
    """fc[i] / mx.nd.norm(fc[i]) for every row (no epsilon; zero row -> NaN)."""
    x = np.asarray(x, dtype=dtype)
    nrm = np.sqrt((x * x).sum(axis=1, keepdims=True, dtype=dtype))
    with np.errstate(divide="ignore", invalid="ignore"):
        return (x / nrm).astype(dtype), nrm.reshape(-1).astype(dtype)


def l2norm_bwd(x, dy, dtype=F32):
    """dX = (dY - y (y . dY)) / |x|   (north_star: per-row L2 with backward)."""
    y, nrm = l2norm_fwd(x, dtype)
    dy = np.asarray(dy, dtype=dtype)
    return ((dy - y * (y * dy).sum(axis=1, keepdims=True, dtype=dtype)) / nrm[:, None]).astype(dtype)


def frob_fwd(x, dtype=F32):
    """anc / mx.nd.norm(anc): whole-block Frobenius scaling (train_efm.py:241)."""
    x = np.asarray(x, dtype=dtype)
    nrm = norm(x, dtype)[0]
    return (x / nrm).astype(dtype), nrm


def frob_bwd(x, dy, dtype=F32):
    y, nrm = frob_fwd(x, dtype)
    dy = np.asarray(dy, dtype=dtype)
    return ((dy - y * (y * dy).sum(dtype=dtype)) / nrm).astype(dtype)


# --------------------------------------------------------------------------------------
# a6  EFM-3 / MFM-2   (lightcnn.py:22-27,32-37,53-58,61-66,123-128; efm_symbol.py:63-77)
# --------------------------------------------------------------------------------------
def mx_maximum(a, b):
    """[MXNet] mshadow_op::maximum: IsNan(a) ? a : (a > b ? a : b)  (so NaN in either
    operand propagates; max(+0,-0) with a=+0 returns b=-0)."""
    a = np.asarray(a); b = np.asarray(b)
    return np.where(np.isnan(a), a, np.where(a > b, a, b))


def mx_minimum(a, b):
    """[MXNet] mshadow_op::minimum: IsNan(a) ? a : (a < b ? a : b)."""
    a = np.asarray(a); b = np.asarray(b)
    return np.where(np.isnan(a), a, np.where(a < b, a, b))


def _thirds(x):
    C = x.shape[1]
    if C % 3:
        raise ValueError("SliceChannel(num_outputs=3) needs C % 3 == 0")
    t = C // 3
    return x[:, 0:t], x[:, t:2 * t], x[:, 2 * t:3 * t]


def efm3_fwd(x):
    """SliceChannel(3, axis=1) -> maximum(maximum(s0,s1),s2), minimum(minimum(s0,s1),s2)
    -> Concat  (lightcnn.py:32-37).  Output has 2C/3 channels."""
    x = np.asarray(x)
    s0, s1, s2 = _thirds(x)
    mx2 = mx_maximum(mx_maximum(s0, s1), s2)
    mn2 = mx_minimum(mx_minimum(s0, s1), s2)
    return np.concatenate([mx2, mn2], axis=1).astype(x.dtype)


def efm3_bwd(x, dy):
    """[MXNet] maximum backward: lhs gets grad where lhs >= rhs, rhs where lhs < rhs;
    minimum backward: lhs where lhs <= rhs, rhs where lhs > rhs.  With the operand order
    of lightcnn.py:33-36 ties resolve to the lowest slice index; NaN comparisons are all
    false so NaN lanes pass no gradient."""
    x = np.asarray(x); dy = np.asarray(dy)
    s0, s1, s2 = _thirds(x)
    t = x.shape[1] // 3
    gmax, gmin = dy[:, :t], dy[:, t:]
    m1 = mx_maximum(s0, s1); n1 = mx_minimum(s0, s1)
    # max branch
    g_m1 = np.where(m1 >= s2, gmax, 0); g2 = np.where(m1 < s2, gmax, 0)
    g0 = np.where(s0 >= s1, g_m1, 0); g1 = np.where(s0 < s1, g_m1, 0)
    # min branch
    h_n1 = np.where(n1 <= s2, gmin, 0); h2 = np.where(n1 > s2, gmin, 0)
    h0 = np.where(s0 <= s1, h_n1, 0); h1 = np.where(s0 > s1, h_n1, 0)
    return np.concatenate([g0 + h0, g1 + h1, g2 + h2], axis=1).astype(x.dtype)


def mfm2_fwd(x):
    """SliceChannel(2) -> maximum(s0, s1)  (efm_symbol.py:63-64,76-77).  C/2 channels."""
    x = np.asarray(x)
    C = x.shape[1]
    if C % 2:
        raise ValueError("SliceChannel(num_outputs=2) needs C % 2 == 0")
    return mx_maximum(x[:, :C // 2], x[:, C // 2:]).astype(x.dtype)


def mfm2_bwd(x, dy):
    x = np.asarray(x); dy = np.asarray(dy)
    C = x.shape[1]
    s0, s1 = x[:, :C // 2], x[:, C // 2:]
    return np.concatenate([np.where(s0 >= s1, dy, 0), np.where(s0 < s1, dy, 0)], axis=1).astype(x.dtype)


# --------------------------------------------------------------------------------------
# a4  cosine_dist (train_efm.py:26-34) and its CSV (train_efm.py:250-255)
# --------------------------------------------------------------------------------------
def cosine_pairs(a, b, dtype=F32):
    """dot(a_i, b_i) / (|a_i| * |b_i|) per row."""
    a = np.asarray(a, dtype=dtype); b = np.asarray(b, dtype=dtype)
    dot = (a * b).sum(axis=1, dtype=dtype)
    na = np.sqrt((a * a).sum(axis=1, dtype=dtype)); nb = np.sqrt((b * b).sum(axis=1, dtype=dtype))
    with np.errstate(divide="ignore", invalid="ignore"):
        return (dot / (na * nb)).astype(dtype)


def cosine_dist_literal(anc, pos, neg, batch_size):
    """Row-by-row restatement returning two Python lists of shape-(1,) arrays exactly like
    the reference (used for the reference-literal CPU baseline)."""
    pos_dist, neg_dist = [], []
    for i in range(batch_size):
        ai, pi, ni = (np.asarray(v[i], dtype=F32) for v in (anc, pos, neg))
        pos_dist.append((np.dot(ai, pi) / (norm(ai) * norm(pi))).astype(F32).reshape(1))
        neg_dist.append((np.dot(ai, ni) / (norm(ai) * norm(ni))).astype(F32).reshape(1))
    return pos_dist, neg_dist


def cosine_allpairs(A, Bm, dtype=F32):
    A = np.asarray(A, dtype=dtype); Bm = np.asarray(Bm, dtype=dtype)
    na = np.sqrt((A * A).sum(axis=1, dtype=dtype)); nb = np.sqrt((Bm * Bm).sum(axis=1, dtype=dtype))
    return ((A @ Bm.T) / (na[:, None] * nb[None, :])).astype(dtype)


def format_cosine_csv(pos, neg):
    """csv.writer(delimiter=' ').writerow([pos.asscalar(), neg.asscalar()]) per row (train_efm.py:250-255): asscalar()
    yields numpy.float32 scalars, which csv stringifies with str() -> shortest float32 repr; '\r\n' line terminator
    (csv default).  Pinned against the reference block itself (tests/golden/ref_python.npz)."""
    rows = []
    for p, n in zip(np.asarray(pos, dtype=F32).reshape(-1), np.asarray(neg, dtype=F32).reshape(-1)):
        rows.append("{} {}\r\n".format(str(F32(p)), str(F32(n))))
    return "".join(rows)


# --------------------------------------------------------------------------------------
# a5  draw_cos_dis_real.py:4-34 (statistics only; plotting is out of scope)
# --------------------------------------------------------------------------------------
def read_cosine_csv(lines, desire_epoch=1):
    """draw_cos_dis_real.py:9-21: keep rows with i > len - int(len/desire_epoch)."""
    pos, neg = [], []
    n = len(lines)
    for i, line in enumerate(lines):
        if i > (n - int(n / desire_epoch)):
            parts = line.split(' ')
            pos.append(float(parts[0])); neg.append(float(parts[1]))
    return pos, neg


def calc_counts(dis_list):
    """draw_cos_dis_real.py:4-7 (np.float -> float)."""
    row = np.array(dis_list).astype(float)
    return np.histogram(row, bins=100)


def calc_lines(pos_dis, neg_dis):
    """draw_cos_dis_real.py:23-34."""
    pc, pb = calc_counts(pos_dis); nc, nb = calc_counts(neg_dis)
    pl = pc / sum(pc); nl = nc / sum(nc)
    return pl, nl, np.cumsum(pl), 1 - np.cumsum(nl), pb, nb


# --------------------------------------------------------------------------------------
# a8  simd_dot (feature_extraction/c_version/Feature.hpp:273-293)
# --------------------------------------------------------------------------------------
def simd_dot(x, y, length=None):
    """4-lane fp32 accumulate over quads while i + 4 < len (STRICT: the last full quad is
    left to the tail), horizontal ((t0+t1)+t2)+t3, then scalar tail.  mul and add round
    separately (no FMA: _mm_add_ps(acc, _mm_mul_ps(X, Y)); x86-64 baseline has no FMA)."""
    x = np.asarray(x, dtype=F32).reshape(-1); y = np.asarray(y, dtype=F32).reshape(-1)
    n = len(x) if length is None else int(length)
    acc = np.zeros(4, dtype=F32)
    i = 0
    while i + 4 < n:
        acc = (acc + (x[i:i + 4] * y[i:i + 4]).astype(F32)).astype(F32)
        i += 4
    s = F32(F32(F32(acc[0] + acc[1]) + acc[2]) + acc[3])
    while i < n:
        s = F32(s + F32(x[i] * y[i]))
        i += 1
    return F32(s)


def simd_dot_rows(G, q):
    """simd_dot(G[r], q) for every row r, vectorised over rows, identical op order."""
    G = np.asarray(G, dtype=F32); q = np.asarray(q, dtype=F32).reshape(-1)
    N, n = G.shape
    acc = np.zeros((N, 4), dtype=F32)
    i = 0
    while i + 4 < n:
        acc = acc + G[:, i:i + 4] * q[None, i:i + 4]
        i += 4
    s = ((acc[:, 0] + acc[:, 1]) + acc[:, 2]) + acc[:, 3]
    while i < n:
        s = s + G[:, i] * q[i]
        i += 1
    return s.astype(F32)


def gallery_scores(G, gnorm, q):
    """sim_r = simd_dot(g_r, q) / (gnorm_r * sqrt(simd_dot(q, q)))  (Feature.hpp:346,356)."""
    q = np.asarray(q, dtype=F32).reshape(-1)
    qn = F32(np.sqrt(simd_dot(q, q)))
    with np.errstate(divide="ignore", invalid="ignore"):
        return (simd_dot_rows(G, q) / (np.asarray(gnorm, dtype=F32) * qn)).astype(F32)


# --------------------------------------------------------------------------------------
# a9  Compare_Face_From_DB(FVV&, sim_th, fv)   (Feature.hpp:345-392), serial semantics
# --------------------------------------------------------------------------------------
def gallery_argmax(G, gnorm, q, sim_th):
    """NaN skipped, sim < th skipped, strict sim > maxsim with maxsim initialised 0 (so only
    positive similarities can win), lowest index on ties; -1 when nothing qualifies."""
    sims = gallery_scores(G, gnorm, q)
    maxsim, rt = F32(0), -1
    for xi, sim in enumerate(sims):
        if np.isnan(sim) or sim < F32(sim_th):
            continue
        if sim > maxsim:
            maxsim, rt = sim, xi
    return rt, maxsim


# --------------------------------------------------------------------------------------
# a10 Compare_Face_From_DB(Configs, fv, size, Face_Maps, K)   (Feature.hpp:295-343)
# --------------------------------------------------------------------------------------
def gallery_topk(G, gnorm, q, K, sim_th):
    """Streaming top-K: fill to K, then replace the current minimum only when min < sim
    (strict, earlier row survives ties).  Returns (status, [(sim, idx)] sorted by
    (sim desc, idx asc)); status -1 and the partial list when a NaN similarity is met
    (the reference returns -1 there).  The list comparator (strict_weak_ordering) lives in
    a file missing from the reference; ascending-by-similarity is implied by the at(0)
    usage at :331-335."""
    sims = gallery_scores(G, gnorm, q)
    keep = []                                    # (sim, idx)
    status = 0
    for xi, sim in enumerate(sims):
        if np.isnan(sim):
            status = -1
            break
        if sim < F32(sim_th):
            continue
        if len(keep) < K:
            keep.append((sim, xi))
        else:
            keep.sort(key=lambda t: (t[0], -t[1]))       # at(0) = minimum (latest among ties)
            if keep[0][0] < sim:
                keep[0] = (sim, xi)
    keep.sort(key=lambda t: (-t[0], t[1]))
    return status, keep


# --------------------------------------------------------------------------------------
# a11 Compare_Face_Person / Compare_Face_DB   (Feature.hpp:763-804)
# --------------------------------------------------------------------------------------
def compare_face_person(G, gnorm, q, sim_th, maxsim, qnorm=None):
    """Feature.hpp:763-790 serial: returns (is_this_person, maxsim')."""
    q = np.asarray(q, dtype=F32).reshape(-1)
    qn = F32(np.sqrt(simd_dot(q, q))) if qnorm is None else F32(qnorm)
    hit = False
    if len(G) == 0:
        return hit, F32(maxsim)
    with np.errstate(divide="ignore", invalid="ignore"):
        sims = (simd_dot_rows(G, q) / (np.asarray(gnorm, dtype=F32) * qn)).astype(F32)
    maxsim = F32(maxsim)
    for sim in sims:
        if np.isnan(sim):
            continue
        if sim >= F32(sim_th) and sim > maxsim:
            maxsim, hit = sim, True
    return hit, maxsim


def compare_face_db(G, gnorm, person_ids, q, sim_th):
    """Feature.hpp:792-804: persons visited in ascending pid (std::map order), running
    maxsim carried across persons; returns the pid of the last person that raised maxsim
    (== identity of the global best row with sim >= th and sim > 0) or -1."""
    person_ids = np.asarray(person_ids)
    q = np.asarray(q, dtype=F32).reshape(-1)
    qn = F32(np.sqrt(simd_dot(q, q)))
    maxsim, best = F32(0), -1
    for pid in np.unique(person_ids):
        rows = np.nonzero(person_ids == pid)[0]
        hit, maxsim = compare_face_person(np.asarray(G)[rows], np.asarray(gnorm)[rows], q, sim_th, maxsim, qn)
        if hit:
            best = int(pid)
    return best, maxsim


# --------------------------------------------------------------------------------------
# a7  feature export text formats (extract_feacture_v2.py:66-79)
# --------------------------------------------------------------------------------------
def format_feature_rows(fc):
    """Each normalised row -> '{},' per element (Python float repr of the fp32 value,
    trailing comma) then '\\n'."""
    y, _ = l2norm_fwd(fc)
    out = []
    for row in y:
        out.append("".join("{},".format(float(e)) for e in row.tolist()) + "\n")
    return "".join(out)


def format_label_rows(labels):
    """'{}\\n'.format(label[v].asscalar()) -> numpy.float32 str, e.g. '12.0'."""
    return "".join("{}\n".format(F32(v)) for v in np.asarray(labels).reshape(-1))


# --------------------------------------------------------------------------------------
# LFW verification protocol shape (feature_extraction/facenet_version/facenet.py:412-473)
# --------------------------------------------------------------------------------------
def lfw_distance(e1, e2, distance_metric=0):
    """facenet.py:412-426."""
    e1 = np.asarray(e1); e2 = np.asarray(e2)
    if distance_metric == 0:
        diff = np.subtract(e1, e2)
        return np.sum(np.square(diff), 1)
    if distance_metric == 1:
        dot = np.sum(np.multiply(e1, e2), axis=1)
        nrm = np.linalg.norm(e1, axis=1) * np.linalg.norm(e2, axis=1)
        return np.arccos(dot / nrm) / math.pi
    raise ValueError("Undefined distance metric %d" % distance_metric)


def lfw_calculate_accuracy(threshold, dist, actual_issame):
    """facenet.py:463-473."""
    predict = np.less(dist, threshold)
    tp = np.sum(np.logical_and(predict, actual_issame))
    fp = np.sum(np.logical_and(predict, np.logical_not(actual_issame)))
    tn = np.sum(np.logical_and(np.logical_not(predict), np.logical_not(actual_issame)))
    fn = np.sum(np.logical_and(np.logical_not(predict), actual_issame))
    tpr = 0 if (tp + fn == 0) else float(tp) / float(tp + fn)
    fpr = 0 if (fp + tn == 0) else float(fp) / float(fp + tn)
    acc = float(tp + tn) / dist.size
    return tpr, fpr, acc


def kfold_indices(n, n_splits):
    """sklearn KFold(n_splits, shuffle=False): contiguous folds, the first n % n_splits
    folds one longer."""
    sizes = np.full(n_splits, n // n_splits, dtype=int)
    sizes[: n % n_splits] += 1
    idx = np.arange(n)
    cur = 0
    for s in sizes:
        test = idx[cur:cur + s]
        train = np.concatenate([idx[:cur], idx[cur + s:]])
        yield train, test
        cur += s


def lfw_calculate_roc(thresholds, e1, e2, actual_issame, nrof_folds=10, distance_metric=0):
    """facenet.py:428-461 with subtract_mean=False."""
    actual_issame = np.asarray(actual_issame)
    nrof_pairs = min(len(actual_issame), e1.shape[0])
    nt = len(thresholds)
    tprs = np.zeros((nrof_folds, nt)); fprs = np.zeros((nrof_folds, nt)); accuracy = np.zeros(nrof_folds)
    dist = lfw_distance(e1, e2, distance_metric)
    for f, (train, test) in enumerate(kfold_indices(nrof_pairs, nrof_folds)):
        acc_train = np.zeros(nt)
        for t, th in enumerate(thresholds):
            _, _, acc_train[t] = lfw_calculate_accuracy(th, dist[train], actual_issame[train])
        best = np.argmax(acc_train)
        for t, th in enumerate(thresholds):
            tprs[f, t], fprs[f, t], _ = lfw_calculate_accuracy(th, dist[test], actual_issame[test])
        _, _, accuracy[f] = lfw_calculate_accuracy(thresholds[best], dist[test], actual_issame[test])
    return np.mean(tprs, 0), np.mean(fprs, 0), accuracy


# --------------------------------------------------------------------------------------
# reference-literal training-step head (pre-trained_efm_v3.py:199-221), for the CPU baseline
# --------------------------------------------------------------------------------------
def reference_literal_step(Wnx, labels, batch_size, margin, rng):
    """anc/pos halves, per-row rejection-sampled negatives copied row by row, explicit
    triplet loss + backward, per-row cosine_dist.  Returns (loss, da, dp, pos_cos, neg_cos, idx)."""
    anc = Wnx[0:batch_size]; pos = Wnx[batch_size:2 * batch_size]
    idx = sample_negatives(labels, batch_size, "train", rng)
    neg = np.stack([np.array(Wnx[j], dtype=F32) for j in idx])          # .asnumpy() row copies
    loss = triplet_loss_fwd(anc, pos, neg, margin)
    da, dp, _ = triplet_loss_bwd(anc, pos, neg, margin)                 # neg is detached (:206-207)
    pc, nc = cosine_dist_literal(anc, pos, neg, batch_size)
    return loss, da, dp, pc, nc, idx
