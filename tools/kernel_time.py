"""Times the mining kernel alone (C2 shape) under the EFM_GRAM_DEBUG / EFM_GRAM_MH profiling switches."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from improving_face_recognition_performance_using_triplet_loss_b200 import _lib as L
B, D = int(os.environ.get("KB_B", 8192)), 256
E, y = bench.make_batch(B, D, 0)
Et = torch.from_numpy(E).cuda().to(torch.bfloat16); yt = torch.from_numpy(y).cuda()
flags = L.BH_NORMALIZE
ws = L.workspace(L.lib.efm_bh_workspace_bytes(B, B, D, L.BF16, flags), Et.device)
loss = torch.empty(B, device="cuda"); ip = torch.empty(B, dtype=torch.int32, device="cuda"); inn = torch.empty_like(ip)
dE = torch.empty(B, D, device="cuda")
L.check(L.lib.efm_bh_triplet_fwd_bwd(L.ptr(Et), L.ptr(yt), B, 0, B, D, L.BF16, 0.2, None, flags, L.ptr(loss), L.ptr(ip), L.ptr(inn), L.ptr(dE), L.ptr(ws), ws.numel(), L.stream()))
torch.cuda.synchronize()
ts = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(40):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()                       # keeps the GPU busy while the host enqueues the timed launch; flushes L2
    s.record(); L.check(L.lib.efm_bh_mine(L.ptr(Et), B, 0, B, D, L.BF16, flags, L.ptr(ws), ws.numel(), L.stream())); e.record()
    torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ts = sorted(ts[5:])
print(json.dumps({"debug": os.environ.get("EFM_GRAM_DEBUG", "0"), "mh": os.environ.get("EFM_GRAM_MH", "auto"), "B": B,
                  "kernel_us_median": 1e3 * ts[len(ts) // 2], "min": 1e3 * ts[0]}))
