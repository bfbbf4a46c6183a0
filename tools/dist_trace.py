"""In-kernel timeline of the fused multi-GPU step (EFM_GRAM_TRACE) under torchrun: when the exchange warps finish their
pushes, how long the TMA producers wait for peer blocks, per rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/dist_trace.py
"""
import os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
os.environ["EFM_GRAM_TRACE"] = str(trace.data_ptr())
from improving_face_recognition_performance_using_triplet_loss_b200 import distributed as D
G = int(os.environ.get("KB_G", 65536)); B = G // world; Dm = 256
E, y = bench.make_batch(G, Dm, 0)
El = torch.from_numpy(E[rank * B:(rank + 1) * B]).to(torch.bfloat16).to(dev); yl = torch.from_numpy(y[rank * B:(rank + 1) * B]).to(dev)
px = D.PeerExchange(B, Dm, torch.bfloat16, device=dev)
for _ in range(3):
    px.step(El, yl, 0.2)
torch.cuda.synchronize(); dist.barrier()
names = {1: "prologue done", 3: "first acc_full", 6: "last tile epilogue done", 7: "exit", 8: "exchange start", 9: "stage A copied",
         10: "stage B copied", 11: "all labels here", 12: "label table done"}
for it in range(3):
    trace.zero_(); torch.cuda.synchronize(); dist.barrier()
    px.step(El, yl, 0.2)
    torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(-1, 16).astype(np.float64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    if it == 2:
        out = ["rank %d: %d CTAs, B_local %d, global %d, exchange %s" % (rank, len(t), B, G, px.mode)]
        for k in sorted(names):
            col = (t[:, k] - t0) / 1e3
            out.append("  %-26s min %8.1f  median %8.1f  max %8.1f us" % (names[k], col.min(), np.median(col), col.max()))
        out.append("  exchange warp 0: %.1f us median waiting for the owners' announcements" % (np.median(t[:, 15]) / 1e3))
        out.append("  producer peer waits: median %.1f us, max %.1f us per CTA over %d waits" % (np.median(t[:, 13]) / 1e3, t[:, 13].max() / 1e3, int(t[:, 14].max())))
        for r in range(world):
            dist.barrier()
            if r == rank:
                print("\n".join(out), flush=True)
px.close()
dist.destroy_process_group()
