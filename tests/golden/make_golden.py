#!/usr/bin/env python
"""Generates tests/golden/*.npz.

Two kinds of vectors:
  * ref_native.npz   -- outputs of the REFERENCE'S OWN compiled functions (simd_dot,
    Compare_Face_From_DB, Compare_Face_Person, Compare_Face_DB from
    feature_extraction/c_version/Feature.hpp, built by oracle/build_ref.sh into oracle/_ref).
    Needs /root/reference (this container).  These pin the oracle for SURVEY rows a8-a11.
  * kat.npz          -- hand-computable known-answer vectors for the MXNet-side arithmetic
    (TripletLoss, EFM tie/NaN rules, normalisation, histogram).  MXNet is not installable here,
    so these are authored from the reference call sites + MXNet's published op semantics
    (parity UNPINNED by the reference itself; the values below are small enough to verify by hand).

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import native  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def make_ref_native():
    native.build()
    ref = native.Ref()
    if not ref.available:
        raise SystemExit("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(20261018)
    out = {}
    # --- simd_dot at the lengths SURVEY 8c(iii) names
    lens = [1, 3, 4, 5, 8, 9, 16, 128, 256, 342, 512, 684]
    for n in lens:
        x = rng.standard_normal(n).astype(np.float32)
        y = rng.standard_normal(n).astype(np.float32)
        out[f"dot_x_{n}"] = x; out[f"dot_y_{n}"] = y
        out[f"dot_r_{n}"] = np.array([ref.simd_dot(x, y)], np.float32)
    out["dot_lens"] = np.array(lens)
    # --- galleries: D in {342 (reference fvSize), 256, 512}
    cases = []
    for ci, (N, D) in enumerate([(64, 342), (257, 256), (1000, 512), (33, 8)]):
        G = rng.standard_normal((N, D)).astype(np.float32)
        ids = (np.arange(N) % max(2, N // 5)).astype(np.int32)
        # cluster structure: rows of an identity share a centroid
        cent = rng.standard_normal((ids.max() + 1, D)).astype(np.float32)
        G = (cent[ids] + 0.3 * G).astype(np.float32)
        # deliberate edge cases: duplicate rows (ties -> lowest index), a zero row (NaN skip)
        G[7] = G[3]
        if N > 40:
            G[40] = 0.0
        gnorm = np.array([np.sqrt(ref.simd_dot(g, g)) for g in G], np.float32)
        Q = []
        Q.append((G[3] + 0.01 * rng.standard_normal(D)).astype(np.float32))   # near rows 3 == 7
        Q.append(G[3].copy())                                                  # exact duplicate tie
        Q.append((-G[5]).astype(np.float32))                                   # mostly negative sims
        Q.append(rng.standard_normal(D).astype(np.float32))                    # unrelated
        Q.append((G[N - 1] + 0.05 * rng.standard_normal(D)).astype(np.float32))
        Q = np.stack(Q)
        ths = np.array([0.0, 0.3, 0.9, 1.5, -1.0], np.float32)
        res_idx = np.zeros((len(Q), len(ths)), np.int32)
        res_pid = np.zeros((len(Q), len(ths)), np.int32)
        res_person_hit = np.zeros((len(Q), len(ths)), np.int32)
        res_person_max = np.zeros((len(Q), len(ths)), np.float32)
        for qi, q in enumerate(Q):
            for ti, th in enumerate(ths):
                res_idx[qi, ti] = ref.compare_face_from_db(G, gnorm, q, float(th))
                res_pid[qi, ti] = ref.compare_face_db(G, gnorm, ids, q, float(th))
                hit, ms = ref.compare_face_person(G[ids == 1], gnorm[ids == 1], q, float(th), 0.0)
                res_person_hit[qi, ti] = int(hit); res_person_max[qi, ti] = ms
        out[f"g{ci}_G"] = G; out[f"g{ci}_gnorm"] = gnorm; out[f"g{ci}_ids"] = ids
        out[f"g{ci}_Q"] = Q; out[f"g{ci}_ths"] = ths
        out[f"g{ci}_idx"] = res_idx; out[f"g{ci}_pid"] = res_pid
        out[f"g{ci}_person_hit"] = res_person_hit; out[f"g{ci}_person_max"] = res_person_max
        cases.append(ci)
    out["gallery_cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "ref_native.npz"), **out)
    print("wrote ref_native.npz with", len(out), "arrays")


def make_kat():
    out = {}
    # (i) triplet 4x3, margin 0.5: rows = active, exactly-zero hinge, inactive, active
    a = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 1]], np.float32)
    p = np.array([[1, 0, 0], [1, 1, 0], [0, 1, 0], [1, 1, 3]], np.float32)
    n = np.array([[1, 1, 0], [2, 0.5, 0.5], [3, 1, 0], [1, 2, 1]], np.float32)
    # d_ap = [1, 1, 0, 4]; d_an = [2, 1.5, 9, 1]; s = d_ap - d_an + 0.5 = [-0.5, 0, -8.5, 3.5]
    out["tl_a"], out["tl_p"], out["tl_n"] = a, p, n
    out["tl_margin"] = np.array([0.5], np.float32)
    out["tl_loss"] = np.array([0, 0, 0, 3.5], np.float32)
    da = np.zeros_like(a); dp = np.zeros_like(a); dn = np.zeros_like(a)
    da[3] = 2 * (n[3] - p[3]); dp[3] = 2 * (p[3] - a[3]); dn[3] = 2 * (a[3] - n[3])
    out["tl_da"], out["tl_dp"], out["tl_dn"] = da, dp, dn
    # margin 1.0 -> s = [0, 0.5, -8, 4]: row 0 is the exactly-zero hinge (relu'(0) = 0), rows 1,3 active
    out["tl_loss_m1"] = np.array([0, 0.5, 0, 4], np.float32)
    # (ii) EFM-3 on (1, 3, 1, 8): ties across slices, NaN lanes, signed zeros
    nan = np.float32(np.nan)
    s0 = np.array([1, 2, 2, 5, nan, 1, +0.0, -1], np.float32)
    s1 = np.array([2, 2, 1, 5, 1, nan, -0.0, -1], np.float32)
    s2 = np.array([3, 1, 2, 5, 1, 1, +0.0, -2], np.float32)
    x = np.stack([s0, s1, s2]).reshape(1, 3, 1, 8)
    out["efm_x"] = x
    # max: [3,2,2,5,nan,nan,+0 (m1 = -0 since a>b false -> b; then -0 > +0 false -> s2 = +0), -1]
    out["efm_max"] = np.array([3, 2, 2, 5, nan, nan, 0.0, -1], np.float32)
    # min: [1,1,1,5,nan,nan, m1 = (+0 < -0 false -> b = -0); (-0 < +0 false -> s2 = +0) -> +0, -2]
    out["efm_min"] = np.array([1, 1, 1, 5, nan, nan, 0.0, -2], np.float32)
    # backward with dy = ones: d(max) routing: col0 -> s2; col1 -> s0 (s0>=s1, m1>=s2); col2 -> s0
    # (m1=2>=s2=2); col3 -> s0; col4 NaN: m1 = NaN, comparisons false -> nothing; col5: m1 = NaN -> nothing;
    # col6: s0>=s1 (0>=-0 true) -> s0 and m1(-0... value) >= s2 true -> s0; col7: s0>=s1 -> s0, m1=-1>=-2 -> s0
    out["efm_dmax_to"] = np.array([2, 0, 0, 0, -1, -1, 0, 0], np.int32)
    # d(min): col0 -> s0 (s0<=s1, n1=1<=3); col1: n1 = min(2,2)=2 (s0<=s1 -> s0), n1<=s2? 2<=1 false -> s2;
    # col2: n1 = 1 (s0<=s1? 2<=1 false -> s1), 1<=2 -> s1; col3 -> s0; col4,5 NaN -> nothing;
    # col6: s0<=s1 true -> s0, n1 <= s2 true -> s0; col7: n1=-1 (s0), -1 <= -2 false -> s2
    out["efm_dmin_to"] = np.array([0, 2, 1, 0, -1, -1, 0, 2], np.int32)
    # (iii) normalisation: 3-4-0 row etc.
    xn = np.array([[3, 4, 0], [0, 0, 2], [1, 2, 2]], np.float32)
    out["norm_x"] = xn
    out["norm_y"] = np.array([[0.6, 0.8, 0], [0, 0, 1], [1 / 3, 2 / 3, 2 / 3]], np.float32)
    out["norm_rows"] = np.array([5, 2, 3], np.float32)
    out["norm_frob"] = np.array([np.sqrt(38.0)], np.float32)
    # (iv) batch-hard: 6 points on a line, labels [0,0,0,1,1,2]
    E = np.array([[0.0], [1.0], [3.0], [3.5], [7.0], [20.0]], np.float32)
    y = np.array([0, 0, 0, 1, 1, 2], np.int32)
    # hardest positive (farthest same-label): 0->2, 1->2 (d=4) vs 0 (d=1) -> 2, 2->0, 3->4, 4->3, 5-> none
    # hardest negative (closest other-label): 0->3, 1->3, 2->3, 3->2, 4->2 (d=16) , 5->4
    out["bh_E"], out["bh_y"] = E, y
    out["bh_idx_p"] = np.array([2, 2, 0, 4, 3, -1], np.int32)
    out["bh_idx_n"] = np.array([3, 3, 3, 2, 2, 4], np.int32)
    # loss (no normalise, margin 0.2): d_ap - d_an + m
    #   0: 9 - 12.25 + .2 <0 ; 1: 4 - 6.25 + .2 < 0; 2: 9 - .25 + .2 = 8.95; 3: 12.25 - .25 + .2 = 12.2;
    #   4: 12.25 - 16 + .2 < 0 ; 5: invalid -> 0
    out["bh_loss"] = np.array([0, 0, 8.95, 12.2, 0, 0], np.float32)
    # (v) histogram: 10 values, desire_epoch = 2 keeps rows i > 10 - 5  -> rows 6..9
    out["hist_pos"] = np.linspace(0.5, 0.95, 10).astype(np.float64)
    out["hist_neg"] = np.linspace(-0.2, 0.25, 10).astype(np.float64)
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **out)
    print("wrote kat.npz with", len(out), "arrays")


if __name__ == "__main__":
    make_kat()
    make_ref_native()
