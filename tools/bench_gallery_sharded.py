"""BASELINE config C5: 1M x 512 gallery (78,579 identities) sharded over the ranks of one box, cosine top-5.
Launch with torchrun (one process per GPU).  Every rank owns N/P contiguous rows (bf16), queries are replicated,
local tcgen05 top-k with global indices, NCCL all-gather of the (nq, K) lists, K-way merge on every rank.
Prints one JSON line per query-batch size on rank 0 (device time, max over ranks)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import improving_face_recognition_performance_using_triplet_loss_b200 as P  # noqa: E402
from improving_face_recognition_performance_using_triplet_loss_b200 import distributed as D  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, Dm, K = int(os.environ.get("GAL_N", 1000000)), 512, 5
lo, hi = D.shard_bounds(N, world, rank)
g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
G = torch.randn(hi - lo, Dm, device="cuda", generator=g)
G = G / G.norm(dim=1, keepdim=True)
gal = P.Gallery(G.to(torch.bfloat16), idx_base=lo)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
for nq in (1, 256, 8192):
    gq = torch.Generator(device="cuda"); gq.manual_seed(7)
    src = torch.randint(0, hi - lo, (nq,), device="cuda", generator=gq)
    Q = G[src] + 0.1 * torch.randn(nq, Dm, device="cuda", generator=gq)
    if world > 1:
        dist.broadcast(Q, 0)                      # replicated queries (rank 0's perturbed rows)
    Qb = Q.to(torch.bfloat16)

    def step():
        s, i = gal.topk(Qb, K, 0.0)
        return D.sharded_topk(s, i) if world > 1 else (s, i)

    for _ in range(3):
        s, i = step()
    ts = []
    for _ in range(8):
        flush.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s, i = step(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor([float(np.median(ts))], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = float(t) * 1e-3
    if rank == 0:
        hit = float((i[:, 0].cpu() == (src.cpu() + lo)).float().mean())        # rank 0's source rows are the true top-1
        n_s = hi - lo
        bound = max(2.0 * nq * n_s * Dm / (pk["bf16_tflops"] * 1e12), n_s * Dm * 2 / (pk["hbm_gbs"] * 1e9))
        print(json.dumps({"path": "gallery_top5_sharded", "config": "C5", "N": N, "D": Dm, "n_gpus": world, "nq": nq, "K": K,
                          "ms": t * 1e3, "queries_per_s": nq / t, "frac_of_per_gpu_bound": bound / t, "top1_is_source": hit}), flush=True)
if world > 1:
    dist.destroy_process_group()
