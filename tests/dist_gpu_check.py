"""Multi-GPU check of the fused exchange step (run under torchrun on N >= 2 GPUs of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_gpu_check.py [--B 1024] [--D 256] [--steps 4] [--time]

Every rank builds the same seeded global batch, keeps its slice, runs `PeerExchange.step` (exchange inside the Gram
kernel, no NCCL) and compares with the single-GPU fused call on the whole batch computed locally: mined indices and loss
bit-equal, gradient rel 1e-5 (multi-GPU == single-GPU on the concatenated batch).  Also checks the NCCL
path (`global_batch_hard`) against the same reference.  Exit code != 0 on any mismatch."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import clustered, rel_err  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1024, help="rows per rank")
    ap.add_argument("--D", type=int, default=256)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--time", action="store_true")
    args = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    import improving_face_recognition_performance_using_triplet_loss_b200 as P
    from improving_face_recognition_performance_using_triplet_loss_b200 import distributed as D

    B, Dm, G = args.B, args.D, args.B * world
    ok = True
    for dtype in (torch.bfloat16, torch.float32):
        px = D.PeerExchange(B, Dm, dtype)
        for s in range(args.steps):
            E, y = clustered(G, Dm, seed=100 + s)
            Eg, yg = torch.from_numpy(E).to(dev).to(dtype), torch.from_numpy(y).to(dev)
            El, yl = Eg[rank * B:(rank + 1) * B].contiguous(), yg[rank * B:(rank + 1) * B].contiguous()
            hg_all = torch.linspace(0.5, 1.5, G, device=dev) if s % 2 else None
            hg = None if hg_all is None else hg_all[rank * B:(rank + 1) * B].contiguous()
            exact = s != 2
            full = P.batch_hard_triplet(Eg, yg, 1.9, mine_bf16=True, head_grad=hg_all)
            sl = slice(rank * B, (rank + 1) * B)
            if exact:
                ref_dE = full[3][sl]
            else:
                ref_dE = P.batch_hard_triplet(Eg, yg, 1.9, mine_bf16=True, row_offset=rank * B, num_local=B, detach_gathered=True,
                                              head_grad=hg)[3][sl]
            l, ip, inn, dE = px.step(El, yl, 1.9, exact_grad=exact, head_grad=hg)
            torch.cuda.synchronize()
            good = (torch.equal(ip, full[1][sl]) and torch.equal(inn, full[2][sl]) and torch.equal(l, full[0][sl]) and
                    rel_err(dE.cpu().numpy(), ref_dE.cpu().numpy()) < 1e-5 and int((full[0] > 0).sum()) > G // 16)
            if not good:
                print("rank %d %s step %d exact=%s MISMATCH: idx_p %d idx_n %d loss %d dE rel %.3g active %d" % (
                    rank, dtype, s, exact, int((ip != full[1][sl]).sum()), int((inn != full[2][sl]).sum()),
                    int((l != full[0][sl]).sum()), rel_err(dE.cpu().numpy(), ref_dE.cpu().numpy()), int((full[0] > 0).sum())), flush=True)
                ok = False
            if s == 0:                                     # the NCCL path on the same data
                l2, ip2, in2, dE2 = D.global_batch_hard(El, yl, 1.9, mine_bf16=True)
                if not (torch.equal(ip2, full[1][sl]) and torch.equal(in2, full[2][sl]) and
                        rel_err(dE2.cpu().numpy(), full[3][sl].cpu().numpy()) < 1e-5):
                    print("rank %d %s NCCL path MISMATCH" % (rank, dtype), flush=True)
                    ok = False
        if args.time and dtype == torch.bfloat16:
            E, y = clustered(G, Dm, seed=7)
            El = torch.from_numpy(E[rank * B:(rank + 1) * B]).to(dev).to(dtype); yl = torch.from_numpy(y[rank * B:(rank + 1) * B]).to(dev)
            for name, fn in (("fused", lambda: px.step(El, yl, 1.9)), ("nccl", lambda: D.global_batch_hard(El, yl, 1.9))):
                for _ in range(3):
                    fn()
                dist.barrier(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    fn()
                e1.record(); torch.cuda.synchronize()
                print("rank %d %s (%s) step %.1f us (B_local=%d, global %d x %d)" % (rank, name, px.mode, e0.elapsed_time(e1) * 100, B, G, Dm), flush=True)
        px.close()
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST CHECK", "PASS" if int(t) else "FAIL", "world", world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) else 1)


if __name__ == "__main__":
    main()
