"""CPU tests: the oracle against the committed golden vectors (reference-native outputs + hand KATs)."""
import random

import numpy as np
import pytest

import oracle as O
from oracle import native


def test_simd_dot_bit_exact_vs_reference(golden):
    g, _ = golden
    port = native.Port()
    for n in g["dot_lens"]:
        x, y, r = g[f"dot_x_{n}"], g[f"dot_y_{n}"], g[f"dot_r_{n}"][0]
        assert O.simd_dot(x, y).tobytes() == r.tobytes(), n
        assert port.simd_dot(x, y).tobytes() == r.tobytes(), n


def test_gallery_argmax_person_db_vs_reference(golden):
    g, _ = golden
    port = native.Port()
    for ci in g["gallery_cases"]:
        G, gn, ids, Q, ths = (g[f"g{ci}_{s}"] for s in ("G", "gnorm", "ids", "Q", "ths"))
        for qi, q in enumerate(Q):
            for ti, th in enumerate(ths):
                r, _ = O.gallery_argmax(G, gn, q, th)
                assert r == g[f"g{ci}_idx"][qi, ti]
                assert port.argmax(G, gn, q, float(th))[0] == r
                pid, _ = O.compare_face_db(G, gn, ids, q, th)
                assert pid == g[f"g{ci}_pid"][qi, ti]
                hit, ms = O.compare_face_person(G[ids == 1], gn[ids == 1], q, th, 0.0)
                assert int(hit) == g[f"g{ci}_person_hit"][qi, ti]
                assert ms.tobytes() == g[f"g{ci}_person_max"][qi, ti].tobytes()


def test_live_reference_if_built(golden):
    """When oracle/_ref exists (this container, or shipped prebuilt) re-run the real reference code."""
    ref = native.Ref()
    if not ref.available:
        pytest.skip("oracle/_ref not built")
    g, _ = golden
    rng = np.random.default_rng(5)
    for D in (7, 342, 512):
        G = rng.standard_normal((50, D)).astype(np.float32)
        gn = np.array([np.sqrt(ref.simd_dot(r, r)) for r in G], np.float32)
        q = (G[11] + 0.1 * rng.standard_normal(D)).astype(np.float32)
        assert ref.compare_face_from_db(G, gn, q, 0.2) == O.gallery_argmax(G, gn, q, 0.2)[0] == 11
        assert ref.simd_dot(G[0], q).tobytes() == O.simd_dot(G[0], q).tobytes()


def test_topk_port_matches_numpy_oracle(golden):
    g, _ = golden
    port = native.Port()
    G, gn, Q = g["g2_G"], g["g2_gnorm"], g["g2_Q"]
    G = G.copy(); gn = gn.copy()
    G[40] = G[41]; gn[40] = gn[41]                      # remove the zero row (NaN -> status -1) for this case
    for q in Q:
        for K in (1, 5, 16):
            st, keep = O.gallery_topk(G, gn, q, K, 0.0)
            st2, s2, i2 = port.topk(G, gn, q, K, 0.0)
            assert st == st2 == 0
            assert [k[1] for k in keep] == [int(v) for v in i2[: len(keep)]]
            assert np.array_equal(np.array([k[0] for k in keep], np.float32), s2[: len(keep)])
    st, _ = O.gallery_topk(g["g2_G"], g["g2_gnorm"], Q[0], 5, 0.0)
    assert st == -1                                      # zero-norm row -> NaN -> "no reg data"


def test_triplet_kat(golden):
    _, k = golden
    assert np.array_equal(O.triplet_loss_fwd(k["tl_a"], k["tl_p"], k["tl_n"], 0.5), k["tl_loss"])
    da, dp, dn = O.triplet_loss_bwd(k["tl_a"], k["tl_p"], k["tl_n"], 0.5)
    assert np.array_equal(da, k["tl_da"]) and np.array_equal(dp, k["tl_dp"]) and np.array_equal(dn, k["tl_dn"])
    assert np.array_equal(O.triplet_loss_fwd(k["tl_a"], k["tl_p"], k["tl_n"], 1.0), k["tl_loss_m1"])
    da, _, _ = O.triplet_loss_bwd(k["tl_a"], k["tl_p"], k["tl_n"], 1.0)
    assert not da[0].any() and da[1].any()               # relu'(0) = 0


def test_triplet_grad_matches_finite_difference():
    rng = np.random.default_rng(0)
    a, p, n = (rng.standard_normal((5, 7)) for _ in range(3))
    da, dp, dn = O.triplet_loss_bwd(a, p, n, 0.3, dtype=np.float64)
    eps = 1e-6
    for arr, g in ((a, da), (p, dp), (n, dn)):
        num = np.zeros_like(arr)
        for idx in np.ndindex(arr.shape):
            arr[idx] += eps; up = O.triplet_loss_fwd(a, p, n, 0.3, np.float64).sum()
            arr[idx] -= 2 * eps; dn_ = O.triplet_loss_fwd(a, p, n, 0.3, np.float64).sum()
            arr[idx] += eps
            num[idx] = (up - dn_) / (2 * eps)
        assert np.allclose(num, g, atol=1e-5)


def test_efm_kat(golden):
    _, k = golden
    y = O.efm3_fwd(k["efm_x"])
    assert np.array_equal(y[0, 0, 0], k["efm_max"], equal_nan=True)
    assert np.array_equal(y[0, 1, 0], k["efm_min"], equal_nan=True)
    for ch, key in ((0, "efm_dmax_to"), (1, "efm_dmin_to")):
        dy = np.zeros((1, 2, 1, 8), np.float32); dy[0, ch] = 1
        dx = O.efm3_bwd(k["efm_x"], dy)[0, :, 0]
        to = np.where(dx.any(0), dx.argmax(0), -1)
        assert np.array_equal(to, k[key])


def test_efm_shapes_and_mfm2():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 6, 3, 3)).astype(np.float32)
    y = O.efm3_fwd(x)
    assert y.shape == (2, 4, 3, 3)
    assert np.array_equal(y[:, :2], np.maximum(np.maximum(x[:, :2], x[:, 2:4]), x[:, 4:]))
    assert np.array_equal(y[:, 2:], np.minimum(np.minimum(x[:, :2], x[:, 2:4]), x[:, 4:]))
    assert np.array_equal(O.mfm2_fwd(x), np.maximum(x[:, :3], x[:, 3:]))
    dy = rng.standard_normal(y.shape).astype(np.float32)
    dx = O.efm3_bwd(x, dy)
    assert np.allclose(dx.reshape(2, 3, 2, 3, 3).sum(1), dy[:, :2] + dy[:, 2:])   # each output's grad lands on one slice
    with pytest.raises(ValueError):
        O.efm3_fwd(np.zeros((1, 4, 2, 2), np.float32))


def test_norm_and_batch_hard_kat(golden):
    _, k = golden
    yn, nr = O.l2norm_fwd(k["norm_x"])
    assert np.allclose(yn, k["norm_y"]) and np.array_equal(nr, k["norm_rows"])
    assert np.allclose(O.norm(k["norm_x"]), k["norm_frob"])
    r = O.batch_hard_triplet(k["bh_E"], k["bh_y"], 0.2, normalize=False)
    assert np.array_equal(r["idx_p"], k["bh_idx_p"]) and np.array_equal(r["idx_n"], k["bh_idx_n"])
    assert np.allclose(r["loss"], k["bh_loss"])


def test_batch_hard_sharded_equals_single():
    """P-way strip mining + summed gradients == single batch."""
    from util import clustered
    E, y = clustered(48, 16, seed=3)
    full = O.batch_hard_triplet(E, y, 0.2, normalize=True)
    dE = np.zeros_like(full["dE_all"]); loss = []; ip = []
    for r in range(4):
        part = O.batch_hard_triplet(E[r * 12:(r + 1) * 12], y[r * 12:(r + 1) * 12], 0.2, normalize=True, E_all=E,
                                    labels_all=y, row_offset=r * 12)
        dE += part["dE_all"]; loss.append(part["loss"]); ip.append(part["idx_p"])
    assert np.allclose(np.concatenate(loss), full["loss"]) and np.array_equal(np.concatenate(ip), full["idx_p"])
    assert np.allclose(dE, full["dE_all"])


def test_batch_hard_fast_port_agrees():
    from util import clustered
    E, y = clustered(64, 32, seed=4)
    slow = O.batch_hard_triplet(E, y, 0.2, normalize=True)
    loss, ip, inn, dE = O.batch_hard_triplet_fast(E, y, 0.2, normalize=True)
    assert np.array_equal(ip, slow["idx_p"]) and np.array_equal(inn, slow["idx_n"])
    assert np.allclose(loss, slow["loss"], atol=1e-5) and np.allclose(dE, slow["dE_all"], atol=1e-4)


def test_batch_hard_edge_cases():
    E = np.random.default_rng(0).standard_normal((5, 4)).astype(np.float32)
    r = O.batch_hard_triplet(E, np.zeros(5, np.int32), 0.2)          # single identity: no negatives
    assert (r["idx_n"] == -1).all() and (r["loss"] == 0).all() and not r["dE_all"].any()
    r = O.batch_hard_triplet(E, np.arange(5, dtype=np.int32), 0.2)   # all singletons: no positives
    assert (r["idx_p"] == -1).all() and (r["loss"] == 0).all()
    E2 = np.concatenate([E, E[:1]]); y2 = np.array([0, 1, 1, 2, 2, 0], np.int32)
    r = O.batch_hard_triplet(E2, y2, 0.2, normalize=False)            # exact duplicate rows 0 and 5
    assert r["idx_p"][0] == 5 and r["idx_p"][5] == 0
    assert (r["idx_n"] != 5).all()                                     # rows 0 and 5 tie -> lowest index wins


def test_sample_negatives_replay():
    lab = np.array([0, 0, 1, 1, 2, 2, 0, 1] * 2, np.float32)
    a = O.sample_negatives(lab, 8, "train", random.Random(7))
    b = O.sample_negatives(lab, 8, "train", random.Random(7))
    assert np.array_equal(a, b) and all(lab[j] != lab[i] for i, j in enumerate(a)) and a.max() < 8
    v = O.sample_negatives(lab, 8, "val", random.Random(7))
    assert v.max() < 16
    # literal replay of the reference loop with the same stream
    rng = random.Random(7); exp = []
    for i in range(8):
        j = rng.randint(0, 7)
        while int(lab[j]) == int(lab[i]):
            j = rng.randint(0, 7)
        exp.append(j)
    assert list(a) == exp
    with pytest.raises(ValueError):
        O.sample_negatives(np.zeros(8), 8)


def test_histogram_and_csv(golden):
    _, k = golden
    lines = ["{} {}\r\n".format(p, n) for p, n in zip(k["hist_pos"], k["hist_neg"])]
    pos, neg = O.read_cosine_csv(lines, desire_epoch=2)
    assert len(pos) == 4 and pos[0] == k["hist_pos"][6]
    pl, nl, pc, nc, pb, nb = O.calc_lines(pos, neg)
    assert len(pl) == 100 and abs(pl.sum() - 1) < 1e-12 and abs(pc[-1] - 1) < 1e-12 and abs(nc[-1]) < 1e-12
    txt = O.format_cosine_csv(np.float32([0.5, 0.25]), np.float32([0.1, -0.2]))
    assert txt.splitlines()[0].split(" ")[0] == "0.5" and txt.endswith("\r\n")


def test_feature_export_format():
    fc = np.array([[3, 4, 0], [0, 0, 2]], np.float32)
    txt = O.format_feature_rows(fc)
    assert txt.splitlines()[0] == "0.6000000238418579,0.800000011920929,0.0,"
    assert O.format_label_rows(np.float32([12, 3])) == "12.0\n3.0\n"


def test_lfw_protocol_shapes():
    rng = np.random.default_rng(0)
    e1 = rng.standard_normal((100, 8)); e2 = e1 + 0.1 * rng.standard_normal((100, 8))
    e2[50:] = rng.standard_normal((50, 8))
    same = np.arange(100) < 50
    tpr, fpr, acc = O.lfw_calculate_roc(np.arange(0, 4, 0.01), e1, e2, same, nrof_folds=10)
    assert acc.shape == (10,) and acc.mean() > 0.9
    folds = list(O.kfold_indices(103, 10))
    assert [len(t) for _, t in folds] == [11, 11, 11] + [10] * 7


def test_mining_contract_cosine_equals_distance_of_normalised_rows():
    """The tensor-core mining contract (oracle.batch_hard_triplet(mine_bf16=True)): for bf16 inputs the
    cosine of the raw rows orders the candidates of an anchor exactly like the squared distance of the normalised rows, so
    the mined indices equal those of the exact fp64 definition of SURVEY 8 a2' whenever the extrema are separated by more
    than rounding noise (clustered data)."""
    from util import clustered
    for seed, (B, D) in enumerate([(256, 64), (512, 256), (300, 128)]):
        E, y = clustered(B, D, seed=seed)
        Eb = O.bf16_round(E.astype(np.float32)).astype(np.float64)          # what a bf16 tensor holds
        a = O.batch_hard_triplet(Eb, y, 0.2, normalize=True, mine_bf16=True)
        b = O.batch_hard_triplet(Eb, y, 0.2, normalize=True, mine_bf16=False)
        assert np.array_equal(a["idx_p"], b["idx_p"]) and np.array_equal(a["idx_n"], b["idx_n"])
        assert np.allclose(a["loss"], b["loss"], rtol=0, atol=1e-12)
        # strip form (multi-GPU contract): rows [64, 192) against the whole batch
        c = O.batch_hard_triplet(Eb[64:192], y[64:192], 0.2, normalize=True, E_all=Eb, labels_all=y, row_offset=64, mine_bf16=True)
        assert np.array_equal(c["idx_p"], a["idx_p"][64:192]) and np.array_equal(c["idx_n"], a["idx_n"][64:192])
    # duplicates: lowest index wins under both metrics
    E, y = clustered(128, 64, seed=9)
    E[100] = E[3]; E[77] = E[3]
    Eb = O.bf16_round(E.astype(np.float32)).astype(np.float64)
    a = O.batch_hard_triplet(Eb, y, 0.2, normalize=True, mine_bf16=True)
    for i in range(128):
        if a["idx_n"][i] in (77, 100):                                      # only allowed when row 3 itself is not a candidate
            assert y[3] == y[i]
