"""GPU parity tests: every CUDA path (through the C ABI via the Python mirror) against the CPU oracle.
Integer/index/byte results are compared bit-exactly; floating point within the tolerance the north star
states: rel 1e-5 for the fp32 paths, rel 1e-3 for bf16-in / fp32-accumulate paths."""
import random

import numpy as np
import pytest
import torch

import oracle as O
from util import clustered, rel_err

pytestmark = pytest.mark.gpu

TOL32, TOL16 = 1e-5, 1e-3


@pytest.fixture(scope="module")
def P():
    import improving_face_recognition_performance_using_triplet_loss_b200 as pkg
    assert torch.cuda.is_available()
    assert pkg._lib.lib.efm_device_arch() == 100, "expected sm_100 (B200)"
    return pkg


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int16 if t.element_size() == 2 else torch.int32).numpy()


# ------------------------------------------------------------------------------------------------ a6
@pytest.mark.parametrize("shape", [(1, 3, 1, 8), (2, 6, 3, 3), (4, 99, 16, 16), (3, 513), (5, 1026), (2, 261, 8, 8),
                                   (2, 198, 7, 5), (70000, 3, 1, 1)])
def test_efm3_fwd_bwd_fp32_bit_exact(P, shape, golden):
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    x.reshape(-1)[:: max(1, x.size // 7)] = 0.5                       # force ties across slices
    if shape == (1, 3, 1, 8):
        x = golden[1]["efm_x"]
    xt = dev(x).requires_grad_(True)
    y = P.efm3(xt)
    yo = O.efm3_fwd(x)
    assert np.array_equal(bits(y), yo.view(np.int32))
    dy = rng.standard_normal(yo.shape).astype(np.float32)
    y.backward(dev(dy))
    assert np.array_equal(bits(xt.grad), O.efm3_bwd(x, dy).view(np.int32))


def test_efm3_nan_and_signed_zero(P, golden):
    x = golden[1]["efm_x"]
    y = P.efm3(dev(x)).cpu().numpy()
    assert np.array_equal(y[0, 0, 0], golden[1]["efm_max"], equal_nan=True)
    assert np.array_equal(y[0, 1, 0], golden[1]["efm_min"], equal_nan=True)
    assert np.array_equal(y.view(np.int32), O.efm3_fwd(x).view(np.int32))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_efm3_mfm2_half_precisions(P, dtype):
    x = torch.randn(3, 12, 5, 7).to(dtype)
    xf = x.float().numpy()
    y = P.efm3(x.cuda())
    assert np.array_equal(y.float().cpu().numpy(), O.efm3_fwd(xf))     # selection is exact in any dtype
    y2 = P.mfm2(x.cuda())
    assert np.array_equal(y2.float().cpu().numpy(), O.mfm2_fwd(xf))
    xt = x.cuda().requires_grad_(True)
    dy = torch.randn(3, 8, 5, 7).to(dtype)
    P.efm3(xt).backward(dy.cuda())
    ref = torch.from_numpy(O.efm3_bwd(xf, dy.float().numpy())).to(dtype)       # one rounding of g0 + h0
    assert torch.equal(xt.grad.cpu(), ref)


def test_mfm2_fp32_and_errors(P):
    x = np.random.default_rng(0).standard_normal((4, 10, 3, 3)).astype(np.float32)
    xt = dev(x).requires_grad_(True)
    y = P.mfm2(xt)
    assert np.array_equal(y.detach().cpu().numpy(), O.mfm2_fwd(x))
    dy = np.ones_like(O.mfm2_fwd(x))
    y.backward(dev(dy))
    assert np.array_equal(xt.grad.cpu().numpy(), O.mfm2_bwd(x, dy))
    with pytest.raises(ValueError):
        P.efm3(torch.zeros(2, 4, 2, 2).cuda())
    with pytest.raises(ValueError):
        P.efm3(torch.zeros(2, 3, 2, 2))                                 # CPU tensor: no fallback
    assert P.efm3(torch.zeros(0, 3, 2, 2).cuda()).shape == (0, 2, 2, 2)


def test_efm_module_matches_torch_composition(P):
    x = torch.randn(2, 3, 16, 16).cuda()
    blk = P.efm(3, 9, 12, 3, 1, 1, 1).cuda()
    y = blk(x)
    c1 = blk.conv_op_1(x); s = c1.chunk(3, 1)
    e1 = torch.cat([torch.maximum(torch.maximum(s[0], s[1]), s[2]), torch.minimum(torch.minimum(s[0], s[1]), s[2])], 1)
    c2 = blk.conv_op_2(e1); s = c2.chunk(3, 1)
    e2 = torch.cat([torch.maximum(torch.maximum(s[0], s[1]), s[2]), torch.minimum(torch.minimum(s[0], s[1]), s[2])], 1)
    assert torch.equal(y, e2)


# ------------------------------------------------------------------------------------------------ a3
@pytest.mark.parametrize("B,D", [(3, 3), (64, 342), (33, 684), (128, 256), (5, 1)])
def test_l2norm_fwd_bwd(P, B, D, golden):
    rng = np.random.default_rng(B * D)
    x = rng.standard_normal((B, D)).astype(np.float32)
    if (B, D) == (3, 3):
        x = golden[1]["norm_x"]
    xt = dev(x).requires_grad_(True)
    y = P.l2_normalize(xt)
    assert rel_err(y.detach().cpu().numpy(), O.l2norm_fwd(x)[0]) < TOL32
    dy = rng.standard_normal((B, D)).astype(np.float32)
    y.backward(dev(dy))
    assert rel_err(xt.grad.cpu().numpy(), O.l2norm_bwd(x.astype(np.float64), dy, np.float64)) < TOL32
    assert rel_err(P.row_norms(dev(x)).cpu().numpy(), O.l2norm_fwd(x)[1]) < TOL32


def test_l2norm_zero_row_is_nan_and_bf16(P):
    x = torch.zeros(2, 8).cuda(); x[1] = 1
    y = P.l2_normalize(x)
    assert torch.isnan(y[0]).all() and torch.allclose(y[1], torch.full((8,), 8 ** -0.5).cuda())
    xb = torch.randn(16, 256).to(torch.bfloat16)
    yb = P.l2_normalize(xb.cuda()).float().cpu().numpy()
    assert rel_err(yb, O.l2norm_fwd(xb.float().numpy())[0]) < 1e-2      # one bf16 rounding of the output


def test_frobenius_norm_fwd_bwd(P):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((40, 342)).astype(np.float32)
    assert rel_err(P.norm(dev(x)).cpu().numpy(), O.norm(x)) < TOL32
    xt = dev(x).requires_grad_(True)
    y = P.frob_normalize(xt)
    assert rel_err(y.detach().cpu().numpy(), O.frob_fwd(x)[0]) < TOL32
    dy = rng.standard_normal(x.shape).astype(np.float32)
    y.backward(dev(dy))
    assert rel_err(xt.grad.cpu().numpy(), O.frob_bwd(x.astype(np.float64), dy, np.float64)) < TOL32


# ------------------------------------------------------------------------------------------------ a1 / a2
def test_triplet_kat(P, golden):
    k = golden[1]
    a, p, n = (dev(k[s]).requires_grad_(True) for s in ("tl_a", "tl_p", "tl_n"))
    loss = P.TripletLoss(margin=0.5)(a, p, n)
    assert np.array_equal(loss.detach().cpu().numpy(), k["tl_loss"])
    loss.backward(torch.ones_like(loss))
    assert np.array_equal(a.grad.cpu().numpy(), k["tl_da"])
    assert np.array_equal(p.grad.cpu().numpy(), k["tl_dp"])
    assert np.array_equal(n.grad.cpu().numpy(), k["tl_dn"])
    a2 = dev(k["tl_a"]).requires_grad_(True)
    l1 = P.TripletLoss(margin=1.0)(a2, dev(k["tl_p"]), dev(k["tl_n"]))
    assert np.array_equal(l1.detach().cpu().numpy(), k["tl_loss_m1"])
    l1.sum().backward()
    assert not a2.grad[0].any() and a2.grad[1].any()                     # relu'(0) = 0


@pytest.mark.parametrize("B,D,margin", [(64, 684, 0.2), (40, 342, 0.2), (1024, 128, 0.5), (7, 5, 1.0)])
def test_triplet_reference_shapes(P, B, D, margin):
    rng = np.random.default_rng(B + D)
    a, p, n = (rng.standard_normal((B, D)).astype(np.float32) * 0.3 for _ in range(3))
    at, pt = dev(a).requires_grad_(True), dev(p).requires_grad_(True)
    nt = dev(n)                                                          # detached like the reference (.asnumpy() copies)
    loss = P.TripletLoss(margin=margin)(at, pt, nt)
    lo = O.triplet_loss_fwd(a, p, n, margin, np.float64)
    assert rel_err(loss.detach().cpu().numpy(), lo) < TOL32
    w = rng.random(B).astype(np.float32)
    loss.backward(dev(w))
    keep = np.abs(lo) > 1e-4                                             # rows away from the hinge kink
    da, dp, _ = O.triplet_loss_bwd(a, p, n, margin, w, np.float64)
    assert rel_err(at.grad.cpu().numpy()[keep], da[keep]) < TOL32
    assert rel_err(pt.grad.cpu().numpy()[keep], dp[keep]) < TOL32
    assert nt.grad is None


def test_triplet_bf16_and_weights(P):
    a, p, n = (torch.randn(32, 256).to(torch.bfloat16) for _ in range(3))
    loss = P.TripletLoss(margin=0.2, weight=0.5)(a.cuda(), p.cuda(), n.cuda(), sample_weight=torch.full((32,), 2.0).cuda())
    lo = O.triplet_loss_fwd(a.float().numpy(), p.float().numpy(), n.float().numpy(), 0.2, np.float64)
    assert rel_err(loss.cpu().numpy(), lo) < TOL16


def test_sample_negatives_replays_reference(P):
    lab = torch.tensor([0, 0, 1, 1, 2, 2, 0, 1] * 2, dtype=torch.float32).cuda()
    got = P.sample_negatives(lab, 8, "train", random.Random(11)).cpu().numpy()
    assert np.array_equal(got, O.sample_negatives(lab.cpu().numpy(), 8, "train", random.Random(11)))
    got = P.sample_negatives(lab, 8, "val", random.Random(11)).cpu().numpy()
    assert np.array_equal(got, O.sample_negatives(lab.cpu().numpy(), 8, "val", random.Random(11)))


# ------------------------------------------------------------------------------------------------ a4 / a5
def test_cosine_dist_and_pairs(P):
    rng = np.random.default_rng(3)
    a, p, n = (rng.standard_normal((40, 342)).astype(np.float32) for _ in range(3))
    pl, nl = P.cosine_dist(dev(a), dev(p), dev(n), 40)
    assert isinstance(pl, list) and len(pl) == 40 and pl[0].shape == (1,)
    assert rel_err(torch.cat(pl).cpu().numpy(), O.cosine_pairs(a, p)) < TOL32
    assert rel_err(torch.cat(nl).cpu().numpy(), O.cosine_pairs(a, n)) < TOL32
    lit_p, _ = O.cosine_dist_literal(a, p, n, 40)
    assert rel_err(torch.cat(pl).cpu().numpy(), np.concatenate(lit_p)) < TOL32
    E = rng.standard_normal((100, 256)).astype(np.float32)
    ia = rng.integers(0, 100, 60); ib = rng.integers(0, 100, 60)
    got = P.cosine_pairs_indexed(dev(E), dev(ia), dev(ib)).cpu().numpy()
    assert rel_err(got, O.cosine_pairs(E[ia], E[ib])) < TOL32


def test_histogram_matches_numpy(P):
    rng = np.random.default_rng(4)
    x = np.clip(rng.normal(0.6, 0.2, 20000), -1, 1).astype(np.float32)
    x[:3] = [x.min(), x.max(), x.max()]
    cnt, edges = P.calc_counts(torch.from_numpy(x).cuda())
    rc, re = O.calc_counts(x.tolist())
    assert np.array_equal(cnt, rc) and np.array_equal(edges, re)
    pl, nl, pc, nc, pb, nb = P.calc_lines(torch.from_numpy(x).cuda(), torch.from_numpy(-x).cuda())
    opl, onl, opc, onc, opb, onb = O.calc_lines(x.tolist(), (-x).tolist())
    assert np.array_equal(pl, opl) and np.array_equal(nc, onc) and np.array_equal(nb, onb)


@pytest.mark.parametrize("M,N,D", [(5, 7, 8), (130, 300, 256), (257, 129, 342), (64, 1000, 512)])
def test_cosine_allpairs(P, M, N, D):
    rng = np.random.default_rng(M + N)
    A = rng.standard_normal((M, D)).astype(np.float32); Bm = rng.standard_normal((N, D)).astype(np.float32)
    got = P.cosine_allpairs(dev(A, torch.bfloat16), dev(Bm, torch.bfloat16)).cpu().numpy()
    ref = O.cosine_allpairs(O.bf16_round(A), O.bf16_round(Bm), np.float64)
    assert np.abs(got - ref).max() < TOL16


# ------------------------------------------------------------------------------------------------ a2'
def _check_bh(P, E, y, margin, normalize, dtype, mine_bf16, tol, row_offset=0, num_local=None, detach=False):
    Et = dev(E, dtype)
    Ein = Et.float().cpu().numpy()                                       # the values the kernel actually sees
    B_all = E.shape[0]
    B = B_all - row_offset if num_local is None else num_local
    tc = (dtype != torch.float32) or bool(mine_bf16)
    sl = slice(row_offset, row_offset + B)
    if margin is None:                                                   # pick a margin that leaves ~half the hinges active
        r0 = O.batch_hard_triplet(Ein[sl], y[sl], 0.0, normalize=normalize, E_all=Ein, labels_all=y,
                                  row_offset=row_offset, mine_bf16=tc)
        margin = float(-np.median(r0["s"][np.isfinite(r0["s"])]))
    loss, ip, inn, dE = P.batch_hard_triplet(Et, dev(y), margin, normalize=normalize, row_offset=row_offset,
                                             num_local=B, mine_bf16=mine_bf16, detach_gathered=detach)
    ref = O.batch_hard_triplet(Ein[sl], y[sl], margin, normalize=normalize, E_all=Ein, labels_all=y,
                               row_offset=row_offset, mine_bf16=tc)
    assert np.array_equal(ip.cpu().numpy(), ref["idx_p"]), "hardest-positive indices differ"
    assert np.array_equal(inn.cpu().numpy(), ref["idx_n"]), "hardest-negative indices differ"
    scale = max(1.0, np.abs(ref["loss"]).max())
    kink = np.abs(ref["s"]) < 1e-3 * scale                               # rows sitting on the hinge kink
    assert (ref["loss"] > 0).sum() > B // 8, "test data leaves too few active hinges"
    assert np.abs(loss.cpu().numpy() - ref["loss"])[~kink].max() / scale < tol
    if not detach and not kink.any():
        assert rel_err(dE.cpu().numpy(), ref["dE_all"]) < tol * 10
    return loss, ip, inn, dE, ref


def test_bh_kat(P, golden):
    k = golden[1]
    for dtype, mb in ((torch.float32, False), (torch.float32, True), (torch.bfloat16, None)):
        loss, ip, inn, dE = P.batch_hard_triplet(dev(k["bh_E"], dtype), dev(k["bh_y"]), 0.2, normalize=False, mine_bf16=mb)
        assert np.array_equal(ip.cpu().numpy(), k["bh_idx_p"]) and np.array_equal(inn.cpu().numpy(), k["bh_idx_n"])
        assert np.allclose(loss.cpu().numpy(), k["bh_loss"], rtol=1e-5)


@pytest.mark.parametrize("B,D", [(128, 256), (200, 64), (333, 342), (1024, 128), (96, 700)])
@pytest.mark.parametrize("normalize", [True, False])
def test_bh_fp32_path(P, B, D, normalize):
    E, y = clustered(B, D, seed=B + D)
    _check_bh(P, E, y, None, normalize, torch.float32, False, TOL32 * 10)


@pytest.mark.parametrize("B,D", [(128, 256), (256, 256), (384, 64), (200, 64), (333, 342), (1000, 128), (513, 512), (2048, 256)])
@pytest.mark.parametrize("normalize", [True, False])
def test_bh_tensor_core_path_bf16(P, B, D, normalize):
    E, y = clustered(B, D, seed=B + D + 1)
    _check_bh(P, E, y, None, normalize, torch.bfloat16, None, TOL16)


def test_bh_fp32_input_tensor_core_mining(P):
    E, y = clustered(640, 256, seed=9)
    _check_bh(P, E, y, None, True, torch.float32, True, TOL32 * 10)


def test_bh_edge_cases(P):
    E = np.random.default_rng(0).standard_normal((300, 64)).astype(np.float32)
    for dtype in (torch.float32, torch.bfloat16):
        loss, ip, inn, dE = P.batch_hard_triplet(dev(E, dtype), dev(np.zeros(300, np.int32)), 0.2)
        assert (inn == -1).all() and (loss == 0).all() and not dE.any()       # single identity
        loss, ip, inn, dE = P.batch_hard_triplet(dev(E, dtype), dev(np.arange(300, dtype=np.int32)), 0.2)
        assert (ip == -1).all() and (loss == 0).all() and not dE.any()        # all singletons
        # exact duplicates: ties resolve to the lowest index
        E2 = E.copy(); E2[299] = E2[0]; E2[150] = E2[0]
        y2 = (np.arange(300) % 7).astype(np.int32); y2[299] = y2[150] = y2[0]
        loss, ip, inn, dE = P.batch_hard_triplet(dev(E2, dtype), dev(y2), 0.2, normalize=False)
        Ein = dev(E2, dtype).float().cpu().numpy()
        ref = O.batch_hard_triplet(Ein, y2, 0.2, normalize=False, mine_bf16=(dtype != torch.float32))
        assert np.array_equal(ip.cpu().numpy(), ref["idx_p"]) and np.array_equal(inn.cpu().numpy(), ref["idx_n"])
        assert not (inn.cpu().numpy() == 150).any() and not (inn.cpu().numpy() == 299).any()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bh_strip_equals_full_batch(P, dtype):
    """Multi-GPU contract on one device: 4 ranks' strips against the gathered batch == the single batch."""
    E, y = clustered(512, 128, seed=21)
    full = P.batch_hard_triplet(dev(E, dtype), dev(y), 1.9)
    assert (full[0] > 0).sum() > 64
    dE = torch.zeros_like(full[3]); losses = []; ips = []; ins = []
    for r in range(4):
        l, ip, inn, d = P.batch_hard_triplet(dev(E, dtype), dev(y), 1.9, row_offset=r * 128, num_local=128)
        dE += d; losses.append(l); ips.append(ip); ins.append(inn)
    assert torch.equal(torch.cat(ips), full[1]) and torch.equal(torch.cat(ins), full[2])
    assert torch.allclose(torch.cat(losses), full[0], rtol=1e-6, atol=1e-7)
    assert rel_err(dE.cpu().numpy(), full[3].cpu().numpy()) < 1e-5
    # detached mode: only local rows receive gradient
    l, ip, inn, d = P.batch_hard_triplet(dev(E, dtype), dev(y), 1.9, row_offset=128, num_local=128, detach_gathered=True)
    assert not d[:128].any() and not d[256:].any() and d[128:256].any()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bh_row_sliced_gradient_equals_full(P, dtype):
    """Multi-GPU exact backward on one device: per-'rank' forward, 'all-gather' of (idx_p, idx_n, wgt), row-sliced
    gradient per rank == the gradient of the single fused call on the whole batch."""
    E, y = clustered(512, 128, seed=31)
    Et, yt = dev(E, dtype), dev(y)
    loss, ip, inn, dE = P.batch_hard_triplet(Et, yt, 1.9)
    assert (loss > 0).sum() > 64
    w = torch.where(loss > 0, 2.0, 0.0) * torch.linspace(0.5, 1.5, 512).cuda()          # non-trivial head gradient
    full = P.batch_hard_backward(Et, ip, inn, 1.9, w / 2)
    parts = [P.batch_hard_grad_rows(Et, ip, inn, w, row_offset=r * 128, num_local=128) for r in range(4)]
    assert rel_err(torch.cat(parts).cpu().numpy(), full.cpu().numpy()) < 1e-5
    ref = O.batch_hard_triplet(Et.float().cpu().numpy(), y, 1.9, head_grad=(w / 2).cpu().numpy(), mine_bf16=(dtype != torch.float32))
    assert rel_err(torch.cat(parts).cpu().numpy(), ref["dE_all"]) < (1e-4 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("phases", [2, 4])
def test_bh_block_major_column_schedule(P, phases, monkeypatch):
    """The multi-GPU tile schedule (column blocks visited block-major, starting at the anchors' own block) mines the same
    triplets as the plain row-major schedule.  EFM_GRAM_PHASES is the single-GPU test hook for it."""
    E, y = clustered(2048, 256, seed=41)
    for src, dups in ((100, (700, 1900)), (1300, (40, 1999))):       # exact duplicate rows in different column blocks: the mined
        for d in dups:                                               # negative must be the lowest index whatever the visiting order
            E[d] = E[src]
    Et, yt = dev(E, torch.bfloat16), dev(y)
    full = P.batch_hard_triplet(Et, yt, 1.9)
    ref = O.batch_hard_triplet(Et.float().cpu().numpy(), y, 1.9, mine_bf16=True)
    assert np.array_equal(full[1].cpu().numpy(), ref["idx_p"]) and np.array_equal(full[2].cpu().numpy(), ref["idx_n"])
    assert set(full[2].cpu().numpy().tolist()) & {40, 100, 700, 1300}    # the duplicated rows do get mined (ties are exercised)
    monkeypatch.setenv("EFM_GRAM_PHASES", str(phases))
    for r in range(2):
        l, ip, inn, _ = P.batch_hard_triplet(Et, yt, 1.9, row_offset=r * 1024, num_local=1024)
        assert torch.equal(ip, full[1][r * 1024:(r + 1) * 1024]) and torch.equal(inn, full[2][r * 1024:(r + 1) * 1024])
        assert torch.equal(l, full[0][r * 1024:(r + 1) * 1024])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("exact", [True, False])
def test_bh_fused_exchange_step_world1(P, dtype, exact):
    """efm_bh_dist_fwd_bwd with a single rank (the exchange kernel copies the rows into the rank's own region, the
    arrival counters gate the later kernels): same mined triplets, loss and gradient as the single-GPU fused call,
    over several steps (both buffers of the region are used)."""
    from improving_face_recognition_performance_using_triplet_loss_b200.distributed import PeerExchange
    B, D = 512, 128
    px = PeerExchange(B, D, dtype)
    try:
        for s in range(3):
            E, y = clustered(B, D, seed=50 + s)
            Et, yt = dev(E, dtype), dev(y)
            ref = P.batch_hard_triplet(Et, yt, 1.9, mine_bf16=True)
            hg = torch.linspace(0.5, 1.5, B).cuda() if s == 2 else None
            if hg is not None:
                ref = P.batch_hard_triplet(Et, yt, 1.9, mine_bf16=True, head_grad=hg)
            l, ip, inn, dE = px.step(Et, yt, 1.9, exact_grad=exact, head_grad=hg)
            assert (ref[0] > 0).sum() > 64
            assert torch.equal(ip, ref[1]) and torch.equal(inn, ref[2])
            assert torch.equal(l, ref[0])
            assert rel_err(dE.cpu().numpy(), ref[3].cpu().numpy()) < 1e-5
    finally:
        px.close()


def test_bh_fused_exchange_rejects_bad_shapes(P):
    from improving_face_recognition_performance_using_triplet_loss_b200.distributed import PeerExchange
    with pytest.raises(ValueError):
        PeerExchange(200, 128, torch.bfloat16)
    with pytest.raises(ValueError):
        PeerExchange(256, 1024, torch.bfloat16)


def test_bh_fused_exchange_two_gpus():
    """The real thing: 2 ranks under torchrun, exchange over NVLink peer memory (skipped on a single-GPU box)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(here, "dist_gpu_check.py"), "--B", "512", "--steps", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_bh_autograd_module(P):
    E, y = clustered(256, 64, seed=5)
    for red in ("none", "sum", "mean"):
        Et = dev(E).requires_grad_(True)
        crit = P.BatchHardTripletLoss(margin=1.9, normalize=True, reduction=red)
        out = crit(Et, dev(y))
        w = torch.rand(256).cuda() if red == "none" else None
        (out * w).sum().backward() if red == "none" else out.backward()
        ref = O.batch_hard_triplet(E, y, 1.9, normalize=True, head_grad=None if w is None else w.cpu().numpy())
        assert (ref["loss"] > 0).sum() > 32
        scale = 1 / 256 if red == "mean" else 1.0
        assert rel_err(Et.grad.cpu().numpy(), ref["dE_all"] * scale) < 1e-4
        assert np.array_equal(crit.idx_p.cpu().numpy(), ref["idx_p"])


def test_bh_full_size_properties(P):
    """BASELINE config C2 (B=8192, D=256): size-independent properties instead of an O(B^2) fp64 oracle."""
    E, y = clustered(8192, 256, seed=77)
    Et = dev(E, torch.bfloat16)
    loss, ip, inn, dE = P.batch_hard_triplet(Et, dev(y), 1.9)
    assert (loss > 0).sum() > 1024
    ip, inn = ip.cpu().numpy(), inn.cpu().numpy()
    assert (y[ip] == y).all() and (ip != np.arange(8192)).all() and (y[inn] != y).all()
    # the mined pair must be the hardest of ALL candidates of the right kind under the mining metric: cosine of the bf16
    # rows with fp32 norms (d = -cos, larger = farther)
    Xm = Et.float().cpu().numpy().astype(np.float64); nr = np.sqrt((Xm * Xm).sum(1))
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 8192, 256)
    for i in rows:
        d = -(Xm @ Xm[i]) / (nr * nr[i])
        same = y == y[i]; same[i] = False
        assert d[ip[i]] >= d[same].max() - 1e-5 and d[inn[i]] <= d[y != y[i]].min() + 1e-5
        assert ip[i] == np.flatnonzero(same)[np.argmax(d[same])] and inn[i] == np.flatnonzero(y != y[i])[np.argmin(d[y != y[i]])]
    # loss given the indices equals the explicit-triplet operator; gradient rows sum to ~0 (translation invariance
    # of the un-normalised loss is lost after normalisation, but x . dE = 0 holds for every row)
    Xt = P.l2_normalize(Et.float())
    l2 = P.triplet_loss(Xt, Xt[torch.from_numpy(ip).cuda().long()], Xt[torch.from_numpy(inn).cuda().long()], 1.9)
    assert torch.allclose(l2, loss, rtol=1e-4, atol=1e-5)
    assert (Et.float() * dE).sum(1).abs().max() < 1e-3
    # fp32-exact path agrees on the indices for this gap-separated data
    _, ip32, in32, _ = P.batch_hard_triplet(Et.float(), dev(y), 0.2, want_grad=False)
    assert (ip32.cpu().numpy() == ip).mean() > 0.98 and (in32.cpu().numpy() == inn).mean() > 0.98   # bf16 rounding flips near-ties only


# ------------------------------------------------------------------------------------------------ a8-a11
def test_simd_dot_bit_exact(P, golden):
    g = golden[0]
    for n in g["dot_lens"]:
        got = P.simd_dot(dev(g[f"dot_x_{n}"]), dev(g[f"dot_y_{n}"])).cpu().numpy()
        assert got.tobytes() == g[f"dot_r_{n}"][0].tobytes(), n


def test_gallery_exact_vs_reference_golden(P, golden):
    g = golden[0]
    for ci in g["gallery_cases"]:
        G, gn, ids, Q, ths = (g[f"g{ci}_{s}"] for s in ("G", "gnorm", "ids", "Q", "ths"))
        gal = P.Gallery(dev(G), dev(gn), dev(ids))
        assert gal.gnorm.cpu().numpy().tobytes() == gn.tobytes()
        assert P.Gallery(dev(G)).gnorm.cpu().numpy().tobytes() == gn.tobytes()      # stored norm recomputed exactly
        sc = gal.scores(dev(Q)).cpu().numpy()
        for qi, q in enumerate(Q):
            ref = O.gallery_scores(G, gn, q)
            nan = np.isnan(ref)                                           # zero-norm row: NaN sign/payload is platform-specific
            assert np.array_equal(np.isnan(sc[qi]), nan)
            assert np.array_equal(sc[qi][~nan].view(np.int32), ref[~nan].view(np.int32))
        for ti, th in enumerate(ths):
            idx, sim = gal.argmax(dev(Q), float(th))
            assert np.array_equal(idx.cpu().numpy(), g[f"g{ci}_idx"][:, ti])
            pid, _ = gal.identify(dev(Q), float(th))
            assert np.array_equal(pid.cpu().numpy(), g[f"g{ci}_pid"][:, ti])
            assert P.Compare_Face_From_DB(gal, float(th), dev(Q[0])) == g[f"g{ci}_idx"][0, ti]
            assert P.Compare_Face_DB(gal, float(th), dev(Q[0])) == g[f"g{ci}_pid"][0, ti]


@pytest.mark.parametrize("K", [1, 5, 16])
def test_gallery_topk_exact(P, golden, K):
    g = golden[0]
    G, gn, Q = g["g2_G"].copy(), g["g2_gnorm"].copy(), g["g2_Q"]
    gal = P.Gallery(dev(G), dev(gn))
    s, i = gal.topk(dev(Q), K, 0.1)
    for qi, q in enumerate(Q):
        sims = O.gallery_scores(G, gn, q)
        ok = ~np.isnan(sims) & (sims >= np.float32(0.1))
        order = sorted(np.flatnonzero(ok), key=lambda r: (-sims[r], r))[:K]
        got_i = i[qi].cpu().numpy(); got_s = s[qi].cpu().numpy()
        assert list(got_i[: len(order)]) == order and (got_i[len(order):] == -1).all()
        assert np.array_equal(got_s[: len(order)], sims[order])
    lst = P.Compare_Face_From_DB(gal, 0.1, dev(Q[0]), 5)
    assert [b for _, b in lst] == list(i[0, : len(lst)].cpu().numpy()) or K != 5


@pytest.mark.parametrize("N,D,nq,K", [(1000, 512, 5, 5), (5000, 256, 300, 5), (777, 342, 130, 1), (4096, 512, 64, 16)])
def test_gallery_topk_tensor_core(P, N, D, nq, K):
    G, ids = clustered(N, D, ids_per=3, noise=0.4, seed=N + D)
    rng = np.random.default_rng(1)
    src = rng.integers(0, N, nq)
    Q = G[src] + 0.05 * rng.standard_normal((nq, D)).astype(np.float32)
    Gb, Qb = O.bf16_round(G), O.bf16_round(Q)
    gn = np.sqrt((Gb.astype(np.float64) ** 2).sum(1)).astype(np.float32)
    gal = P.Gallery(dev(G, torch.bfloat16), dev(gn), idx_base=1000)
    s, i = gal.topk(dev(Q, torch.bfloat16), K, 0.0)
    qn = np.sqrt((Qb.astype(np.float64) ** 2).sum(1))
    sims = (Qb.astype(np.float64) @ Gb.astype(np.float64).T) / (qn[:, None] * gn[None, :].astype(np.float64))
    i = i.cpu().numpy() - 1000; s = s.cpu().numpy()
    assert (i[:, 0] == src).all()                                          # the perturbed source row is top-1
    for q in range(nq):
        ok = sims[q] >= 0.0
        order = sorted(np.flatnonzero(ok), key=lambda r: (-sims[q, r], r))[:K]
        # scores within tolerance; identity of the k-th entry may swap only between near-equal scores
        assert np.abs(s[q, : len(order)] - sims[q, order]).max() < TOL16
        assert set(i[q, : len(order)]) == set(order) or np.abs(sims[q, i[q, : len(order)]] - sims[q, order]).max() < TOL16


def test_merge_topk_and_shards(P):
    G, _ = clustered(3000, 256, seed=2)
    Q = G[:50] + 0.01
    gn = np.sqrt((O.bf16_round(G).astype(np.float64) ** 2).sum(1)).astype(np.float32)
    whole = P.Gallery(dev(G, torch.bfloat16), dev(gn)).topk(dev(Q, torch.bfloat16), 5, 0.0)
    parts_s, parts_i = [], []
    for r in range(3):
        sh = P.Gallery(dev(G[r * 1000:(r + 1) * 1000], torch.bfloat16), dev(gn[r * 1000:(r + 1) * 1000]), idx_base=r * 1000)
        s, i = sh.topk(dev(Q, torch.bfloat16), 5, 0.0)
        parts_s.append(s); parts_i.append(i)
    ms, mi = P.merge_topk(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(mi, whole[1]) and torch.equal(ms, whole[0])


def test_feature_export(P):
    fc = torch.tensor([[3.0, 4.0, 0.0], [0.0, 0.0, 2.0]]).cuda()
    net = lambda d: (torch.zeros(2, 4).cuda(), fc)                         # noqa: E731
    out, fn = P.extract_features(net, None)
    assert P.format_feature_rows(fn) == O.format_feature_rows(fc.cpu().numpy())
    assert P.format_label_rows(torch.tensor([12.0, 3.0])) == O.format_label_rows(np.float32([12, 3]))


def test_register_db_roundtrip(P, tmp_path):
    """Register_DB_File / Select_Registed_Data (Feature.hpp:393-419,453-489) through the sqlite file: stored norms are the
    exact sqrt(simd_dot), duplicates of rows already in the DB are skipped (sim >= 0.99999), and a gallery loaded from the
    file answers like the in-memory one."""
    F = P.formats
    rng = np.random.default_rng(5)
    G = rng.standard_normal((12, 342)).astype(np.float32)
    paths = ["/faces/p%02d/im%d.jpg" % (i % 5, i) for i in range(12)]
    db = str(tmp_path / "reg.db")
    F.create_register_db(db)
    assert F.register_db_file(db, paths[:8], dev(G[:8])) == 8
    ids, names, G2, gn = F.select_registered(db, 342)
    assert names == ["p%02d" % (i % 5) for i in range(8)] and np.array_equal(G2, G[:8])
    assert np.array_equal(gn, np.array([np.sqrt(np.float32(O.simd_dot(g, g))) for g in G[:8]], dtype=np.float32))
    # second call: rows 6, 7 are already registered -> skipped; 8..11 are new.  Duplicates WITHIN one call are not detected
    # (AllMembers is loaded once before the loop, Feature.hpp:463-464)
    batch = np.concatenate([G[6:12], G[11:12]])
    assert F.register_db_file(db, paths[6:12] + [paths[11]], dev(batch)) == 5
    gal, ids, names = F.load_gallery(db, 342)
    assert gal.N == 13 and ids.tolist() == list(range(1, 14))
    q = (G[9] + 0.01 * rng.standard_normal(342)).astype(np.float32)
    idx, sim = gal.argmax(dev(q).reshape(1, -1), 0.5)
    ref_idx, ref_sim = O.gallery_argmax(np.concatenate([G[:8], G[8:12], G[11:12]]), gal.gnorm.cpu().numpy(), q, 0.5)
    assert int(idx[0]) == ref_idx == 9 and names[9] == "p04"
    assert bits(sim)[0] == np.float32(ref_sim).view(np.int32)


def test_lfw_pair_distances_on_device(P):
    """verification.pair_distances (GPU operators) == facenet.distance on the gathered rows, both metrics, and the ROC built
    from them equals the host protocol on the same embeddings (C4-shaped: index pairs into one embedding matrix)."""
    import os
    V = P.verification
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python.npz"))
    e1, e2, same = g["lfw_e1"].astype(np.float32), g["lfw_e2"].astype(np.float32), g["lfw_issame"]
    E = np.concatenate([e1, e2]); n = e1.shape[0]
    ia, ib = np.arange(n, dtype=np.int32), np.arange(n, 2 * n, dtype=np.int32)
    for m, tol in ((0, 1e-5), (1, 1e-4)):
        d = V.pair_distances(dev(E), ia, ib, distance_metric=m)
        ref = V.distance(e1.astype(np.float64), e2.astype(np.float64), m)
        assert np.abs(d - ref).max() < tol * max(1.0, np.abs(ref).max())
    d0 = V.pair_distances(dev(E), ia, ib, distance_metric=0)
    acc_dev = V.calculate_accuracy(1.0, d0, same)
    assert abs(acc_dev[2] - float(g["lfw_acc_at_1"][2])) <= 2.0 / n          # at most a borderline pair or two flips


def test_batch_hard_step_object_equals_functional_call(P):
    """BatchHardStep (everything allocated once, one library call per step) == batch_hard_triplet, over several batches,
    with and without a head gradient; its buffers are reused."""
    B, D = 512, 128
    st = P.BatchHardStep(B, D, torch.bfloat16, 1.9)
    for s in range(3):
        E, y = clustered(B, D, seed=70 + s)
        Et, yt = dev(E, torch.bfloat16), dev(y)
        hg = torch.linspace(0.5, 1.5, B).cuda() if s == 2 else None
        ref = P.batch_hard_triplet(Et, yt, 1.9, head_grad=hg)
        l, ip, inn, dE = st.step(Et, yt, head_grad=hg)
        assert torch.equal(l, ref[0]) and torch.equal(ip, ref[1]) and torch.equal(inn, ref[2])
        assert rel_err(dE.cpu().numpy(), ref[3].cpu().numpy()) < 1e-5
        assert l.data_ptr() == st.loss.data_ptr()
    with pytest.raises(ValueError):
        st.step(dev(E, torch.float32), yt)
    with pytest.raises(ValueError):
        st.step(Et, yt.long())
