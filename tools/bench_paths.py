"""Secondary-path measurements (not the bench.py headline): EFM-3 over the LightCNN-29 sites (C1), all-pairs cosine
(C4), gallery top-k per shard (C5), row-wise ops.  CUDA events on the current stream, L2 flushed between iterations,
fractions against MEASURED_PEAKS.json.  One JSON line per measurement."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import improving_face_recognition_performance_using_triplet_loss_b200 as P  # noqa: E402

PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)) * 1e-3


def out(**kw):
    print(json.dumps(kw), flush=True)


# LightCNN-29 EFM-3 sites at N=128, 128x128 gray: (count, C_in, H, W)
SITES = [(1, 99, 128, 128), (1, 66, 64, 64), (2, 99, 64, 64), (1, 198, 64, 64), (2, 132, 32, 32), (3, 198, 32, 32),
         (1, 387, 32, 32), (3, 258, 16, 16), (4, 387, 16, 16), (1, 261, 16, 16), (4, 174, 8, 8), (6, 261, 8, 8), (1, 1026, 1, 1)]


def efm_sites(dtype):
    tot_b, tot_t = 0.0, 0.0
    for cnt, C, H, W in SITES:
        x = torch.randn(128, C, H, W, device="cuda", dtype=dtype)
        t = timed(lambda: P.efm3(x), iters=5)
        b = x.numel() * x.element_size() * (1 + 2 / 3)
        tot_b += cnt * b; tot_t += cnt * t
        out(path="efm3_fwd", dtype=str(dtype), shape=[128, C, H, W], ms=t * 1e3, gbs=b / t / 1e9, frac=b / t / 1e9 / PK["hbm_gbs"])
        del x
    out(path="efm3_fwd_all_30_sites", dtype=str(dtype), ms=tot_t * 1e3, bytes=tot_b, gbs=tot_b / tot_t / 1e9,
        frac=tot_b / tot_t / 1e9 / PK["hbm_gbs"])
    x = torch.randn(128, 198, 64, 64, device="cuda", dtype=dtype, requires_grad=True)
    y = P.efm3(x); g = torch.randn_like(y)
    t = timed(lambda: torch.autograd.grad(y, x, g, retain_graph=True), iters=5)
    b = x.numel() * x.element_size() * (2 + 2 / 3)
    out(path="efm3_bwd", dtype=str(dtype), shape=[128, 198, 64, 64], ms=t * 1e3, gbs=b / t / 1e9, frac=b / t / 1e9 / PK["hbm_gbs"])
    x2 = torch.randn(128, 198, 64, 64, device="cuda", dtype=dtype)
    t = timed(lambda: P.mfm2(x2), iters=5)
    b = x2.numel() * x2.element_size() * 1.5
    out(path="mfm2_fwd", dtype=str(dtype), shape=[128, 198, 64, 64], ms=t * 1e3, gbs=b / t / 1e9, frac=b / t / 1e9 / PK["hbm_gbs"])


def efm_sites_graph(dtype):
    """All 30 EFM-3 sites of one LightCNN-29 forward (C1: N=128, 1x128x128 input) back to back as ONE CUDA graph launched
    through the Python API calls (captured once): what the activation costs inside a network forward, without a host
    round trip and an event pair per site.  Inputs are distinct tensors per site (4.8 GB of traffic per replay at fp32)."""
    xs = []
    for cnt, C, H, W in SITES:
        for _ in range(cnt):
            xs.append(torch.randn(128, C, H, W, device="cuda", dtype=dtype))
    tot_b = sum(x.numel() * x.element_size() * (1 + 2 / 3) for x in xs)
    for x in xs[:3]:
        P.efm3(x)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for x in xs:
            P.efm3(x)
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ys = [P.efm3(x) for x in xs]
    t = timed(g.replay, iters=10)
    out(path="efm3_fwd_all_30_sites_one_graph", dtype=str(dtype), launches=len(xs), ms=t * 1e3, bytes=tot_b, gbs=tot_b / t / 1e9,
        frac=tot_b / t / 1e9 / PK["hbm_gbs"])
    del ys, g, xs


def rowops():
    B, D = 1 << 20, 256
    x = torch.randn(B, D, device="cuda")
    t = timed(lambda: P.l2_normalize(x), iters=5)
    out(path="l2norm_fwd", shape=[B, D], ms=t * 1e3, gbs=2 * B * D * 4 / t / 1e9, frac=2 * B * D * 4 / t / 1e9 / PK["hbm_gbs"])
    a, p, n = (torch.randn(B, D, device="cuda") for _ in range(3))
    t = timed(lambda: P.triplet_loss(a, p, n, 0.2), iters=5)
    out(path="triplet_fwd", shape=[B, D], ms=t * 1e3, gbs=3 * B * D * 4 / t / 1e9, frac=3 * B * D * 4 / t / 1e9 / PK["hbm_gbs"])
    t = timed(lambda: P.cosine_pairs(a, p), iters=5)
    out(path="cosine_pairs", shape=[B, D], ms=t * 1e3, gbs=2 * B * D * 4 / t / 1e9, frac=2 * B * D * 4 / t / 1e9 / PK["hbm_gbs"])


def allpairs():
    M, D = 13233, 256
    A = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    t = timed(lambda: P.cosine_allpairs(A), iters=5)
    out(path="cosine_allpairs_C4", shape=[M, M, D], ms=t * 1e3, write_gbs=M * M * 4 / t / 1e9, frac_hbm=M * M * 4 / t / 1e9 / PK["hbm_gbs"],
        tflops=2.0 * M * M * D / t / 1e12)


def gallery():
    N, D = 125000, 512                                            # one C5 shard (1M x 512 over 8 GPUs)
    G = torch.randn(N, D, device="cuda")
    G = G / G.norm(dim=1, keepdim=True)
    gal = P.Gallery(G.to(torch.bfloat16))
    for nq in (1, 256, 8192):
        Q = (G[torch.randint(0, N, (nq,), device="cuda")] + 0.1 * torch.randn(nq, D, device="cuda")).to(torch.bfloat16)
        t = timed(lambda: gal.topk(Q, 5, 0.0), iters=5)
        fl = 2.0 * nq * N * D
        by = N * D * 2
        bound_t = max(fl / (PK["bf16_tflops"] * 1e12), by / (PK["hbm_gbs"] * 1e9))
        out(path="gallery_topk_bf16_shard", N=N, D=D, nq=nq, K=5, ms=t * 1e3, queries_per_s=nq / t, tflops=fl / t / 1e12,
            gbs=by / t / 1e9, frac_of_bound=bound_t / t)
    gal32 = P.Gallery(G)
    for nq in (1, 4):
        Q = G[:nq] + 0.01
        t = timed(lambda: gal32.topk(Q, 5, 0.0), iters=5)
        out(path="gallery_topk_f32_exact_shard", N=N, D=D, nq=nq, K=5, ms=t * 1e3, queries_per_s=nq / t, gbs=N * D * 4 / t / 1e9,
            frac_hbm=N * D * 4 / t / 1e9 / PK["hbm_gbs"])


if __name__ == "__main__":
    which = sys.argv[1:] or ["efm", "rowops", "allpairs", "gallery"]
    if "efm" in which:
        efm_sites(torch.float32); efm_sites(torch.bfloat16)
    if "efmgraph" in which:
        efm_sites_graph(torch.float32); efm_sites_graph(torch.bfloat16)
    if "rowops" in which:
        rowops()
    if "allpairs" in which:
        allpairs()
    if "gallery" in which:
        gallery()
