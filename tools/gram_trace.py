"""Per-CTA phase timestamps of the mining kernel (EFM_GRAM_TRACE): where the fixed cost of a launch goes."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
trace = torch.zeros(148 * 2 * 16, dtype=torch.int64, device="cuda")
os.environ["EFM_GRAM_TRACE"] = str(trace.data_ptr())
from improving_face_recognition_performance_using_triplet_loss_b200 import _lib as L
B, D = int(os.environ.get("KB_B", 8192)), 256
E, y = bench.make_batch(B, D, 0)
Et = torch.from_numpy(E).cuda().to(torch.bfloat16); yt = torch.from_numpy(y).cuda()
flags = L.BH_NORMALIZE
ws = L.workspace(L.lib.efm_bh_workspace_bytes(B, B, D, L.BF16, flags), Et.device)
loss = torch.empty(B, device="cuda"); ip = torch.empty(B, dtype=torch.int32, device="cuda"); inn = torch.empty_like(ip)
dE = torch.empty(B, D, device="cuda")
for _ in range(3):
    L.check(L.lib.efm_bh_triplet_fwd_bwd(L.ptr(Et), L.ptr(yt), B, 0, B, D, L.BF16, 0.2, None, flags, L.ptr(loss), L.ptr(ip), L.ptr(inn), L.ptr(dE), L.ptr(ws), ws.numel(), L.stream()))
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for it in range(5):
    flush.zero_(); trace.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); L.check(L.lib.efm_bh_mine(L.ptr(Et), B, 0, B, D, L.BF16, flags, L.ptr(ws), ws.numel(), L.stream())); e.record()
    torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(-1, 16)[:, :8].astype(np.float64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    rows.append((s.elapsed_time(e) * 1e3, rel))
ev, rel = rows[-1]
names = ["entry", "prologue done", "first loads issued", "first acc_full", "first tile epilogue done", "last MMA committed", "last tile epilogue done", "exit"]
print("event time us", round(ev, 1), "ctas", len(rel))
for k, n in enumerate(names):
    col = rel[:, k]; col = col[col >= 0]
    print(f"{n:28s} min {col.min():7.2f}  median {np.median(col):7.2f}  max {col.max():7.2f} us")
