"""LFW protocol and the a5 statistics against vectors produced by the REFERENCE'S OWN Python functions
(tests/golden/ref_python.npz, see tests/golden/make_golden_py.py): bit-for-bit."""
import os

import numpy as np
import pytest

import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def V():
    from improving_face_recognition_performance_using_triplet_loss_b200 import verification
    return verification


def test_kfold_matches_sklearn(V):
    from sklearn.model_selection import KFold
    for n, k in ((600, 10), (6000, 10), (103, 10), (7, 3)):
        ours = list(V.kfold_indices(n, k))
        ref = list(KFold(n_splits=k, shuffle=False).split(np.arange(n)))
        assert len(ours) == len(ref)
        for (a, b), (c, d) in zip(ours, ref):
            assert np.array_equal(a, c) and np.array_equal(b, d)


def test_distance_accuracy_roc_pinned_to_reference(V, g):
    e1, e2, same = g["lfw_e1"], g["lfw_e2"], g["lfw_issame"]
    for m in (0, 1):
        assert np.array_equal(V.distance(e1, e2, m), g["lfw_dist_m%d" % m])
        thr = g["lfw_thr_m%d" % m]
        for sm in (0, 1):
            tpr, fpr, acc = V.calculate_roc(thr, e1, e2, same, nrof_folds=10, distance_metric=m, subtract_mean=bool(sm))
            tag = "m%d_sm%d" % (m, sm)
            assert np.array_equal(tpr, g["lfw_tpr_" + tag]) and np.array_equal(fpr, g["lfw_fpr_" + tag])
            assert np.array_equal(acc, g["lfw_acc_" + tag])
            if not sm:                                      # the oracle's own restatement (subtract_mean=False form)
                t2, f2, a2 = O.lfw_calculate_roc(thr, e1, e2, same, nrof_folds=10, distance_metric=m)
                assert np.array_equal(t2, tpr) and np.array_equal(f2, fpr) and np.array_equal(a2, acc)
        assert np.array_equal(O.lfw_distance(e1, e2, m), g["lfw_dist_m%d" % m])
    assert np.array_equal(np.array(V.calculate_accuracy(1.0, g["lfw_dist_m0"], same)), g["lfw_acc_at_1"])
    assert np.array_equal(np.array(V.calculate_val_far(1.0, g["lfw_dist_m0"], same)), g["lfw_valfar_at_1"])
    with pytest.raises(ValueError):
        V.distance(e1, e2, 2)


def test_calculate_val_unpinned_but_consistent(V, g):
    """The reference's calculate_val raises on this scipy (recorded in the golden file); ours interpolates over the
    de-duplicated FAR curve: the FAR it reaches on the training folds brackets the target."""
    assert "duplicates" in bytes(g["lfw_val_error"]).decode()
    e1, e2, same = g["lfw_e1"], g["lfw_e2"], g["lfw_issame"]
    val, std, far = V.calculate_val(g["lfw_thr_m0"], e1, e2, same, 1e-2, nrof_folds=10)
    assert 0.0 < val < 1.0 and 0.0 <= far < 0.1 and std >= 0.0
    val2, _, far2 = V.calculate_val(g["lfw_thr_m0"], e1, e2, same, 0.2, nrof_folds=10)
    assert val2 >= val and far2 >= far                     # a looser FAR target accepts more


def test_a5_statistics_pinned_to_reference(g, tmp_path):
    """draw_cos_dis_real.py (row a5): the oracle restatement == the reference's functions run on the same file."""
    p = str(tmp_path / "cosine_similarity.csv")
    with open(p, "wb") as f:
        f.write(bytes(g["a5_csv_text"]))
    lines = open(p).readlines()
    for ep in (1, 5, 7):
        pos, neg = O.read_cosine_csv(lines, desire_epoch=ep)
        assert np.array_equal(np.array(pos), g["a5_read_pos_ep%d" % ep]) and np.array_equal(np.array(neg), g["a5_read_neg_ep%d" % ep])
    pos, neg = O.read_cosine_csv(lines, desire_epoch=5)
    cnt, edges = O.calc_counts(pos)
    assert np.array_equal(cnt, g["a5_counts_pos"]) and np.array_equal(edges, g["a5_edges_pos"])
    got = O.calc_lines(pos, neg)
    for k, v in zip(["pos_line", "neg_line", "pos_cdf", "neg_cdf", "pos_bins", "neg_bins"], got):
        assert np.array_equal(np.asarray(v), g["a5_" + k]), k
    # the package's host-side reader is the same function
    from improving_face_recognition_performance_using_triplet_loss_b200 import cosine
    p2, n2 = cosine.read_csv(p, desire_epoch=5)
    assert p2 == pos and n2 == neg
