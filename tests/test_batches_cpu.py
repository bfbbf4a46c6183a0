"""Triplet batch construction against the reference's own define_pos / DataIter
(train_efm.py:37-114) run in the build container: tests/golden/ref_python.npz."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python.npz")


@pytest.fixture(scope="module")
def Bt():
    from improving_face_recognition_performance_using_triplet_loss_b200 import batches
    return batches


def test_define_pos_and_pair_batches_pinned_to_reference(Bt):
    g = np.load(GOLD)
    imgs, lab = torch.from_numpy(g["f1_imgs"]), torch.from_numpy(g["f1_labels"])
    pos = Bt.define_pos(lab)
    assert list(pos.keys()) == g["f1_pos_keys"].tolist()                       # dict order = order of first appearance
    assert np.array_equal(imgs[list(pos.values())].numpy(), g["f1_pos_imgs"])  # the first sample of every identity
    it = Bt.TripletPairIter(imgs, lab, batch_size=4, length=6, pos=pos)
    data = np.stack([d.numpy() for d, _ in it]); labels = np.stack([l.numpy() for _, l in it])
    assert data.shape == g["f1_batches_data"].shape == (6, 8, 1, 2, 3)
    assert np.array_equal(data, g["f1_batches_data"]) and np.array_equal(labels, g["f1_batches_label"])
    # the contract the loss head relies on (train_efm.py:232-233): rows [0,B) anchors, [B,2B) their positives, labels twice
    for d, l in it:
        assert torch.equal(l[:4], l[4:])
        for r in range(4):
            first = pos[int(l[r])]
            assert torch.equal(d[4 + r], imgs[first])
    assert len(it) == 6 and it.batch_indices(0).shape == (8,)
    with pytest.raises(ValueError):
        Bt.TripletPairIter(imgs, lab, batch_size=4, length=7)
    with pytest.raises(KeyError):
        Bt.TripletPairIter(imgs, lab, batch_size=4, pos={0: 0})


def test_pk_batches_properties(Bt):
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 50, 1000)
    lab[lab == 7] = 8                                                          # identity 7 vanishes
    b = Bt.pk_batches(lab, P=16, K=4, n_batches=20, seed=3)
    assert b.shape == (20, 64) and b.dtype == torch.int64
    for row in b.numpy():
        assert len(set(row.tolist())) == 64                                    # no sample twice
        ids = lab[row].reshape(16, 4)
        assert (ids == ids[:, :1]).all() and len(set(ids[:, 0].tolist())) == 16
    assert torch.equal(b, Bt.pk_batches(lab, 16, 4, 20, seed=3))               # deterministic per seed
    with pytest.raises(ValueError):
        Bt.pk_batches(np.arange(10), P=2, K=2, n_batches=1)


def test_negative_selection_pinned_to_reference():
    """Row a2: the rejection-sampling loops of train_efm.py:234-238 (training, anchor half) and :268-272 (validation, both
    halves), executed from the reference tree with a seeded `random` (tests/golden/make_golden_py.py) == the oracle replay
    == the package's host function, draw for draw."""
    import random

    import oracle as O
    from improving_face_recognition_performance_using_triplet_loss_b200.loss import sample_negatives
    g = np.load(GOLD)
    lab = g["a2_labels"]
    for pool in ("train", "val"):
        for seed in (0, 1, 12345):
            want = g["a2_%s_seed%d" % (pool, seed)]
            assert np.array_equal(O.sample_negatives(lab, 16, pool=pool, rng=random.Random(seed)), want)
            got = sample_negatives(torch.from_numpy(lab), 16, pool=pool, rng=random.Random(seed))
            assert np.array_equal(got.numpy(), want)
            assert (lab[want].astype(int) != lab[:16].astype(int)).all()
