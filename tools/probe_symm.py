"""Probe: does this box support torch symmetric memory with NVLS multicast? (run under torchrun)"""
import os, sys, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    print("rank", rank, "symm ok: buffer_ptrs", [hex(p) for p in h.buffer_ptrs][:3], "multicast_ptr", hex(h.multicast_ptr),
          "signal_pad_ptrs", len(h.signal_pad_ptrs), "backend", getattr(symm, "get_backend", lambda d: "?")(dev) if hasattr(symm, "get_backend") else "?", flush=True)
except Exception as e:
    print("rank", rank, "symm FAILED:", repr(e)[:400], flush=True)
try:
    from cuda import cuda as cu
    err, = cu.cuInit(0)
    err, d = cu.cuDeviceGet(lr)
    err, v = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d)
    err2, f = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, d)
    print("rank", rank, "MULTICAST_SUPPORTED", v, "FABRIC_SUPPORTED", f, flush=True)
except Exception as e:
    print("rank", rank, "cuda-python probe failed", repr(e)[:200], flush=True)
# P2P copy-engine bandwidth rank0 -> rank1 and SM-store bandwidth for reference
if world >= 2:
    n = 64 << 20
    a = torch.empty(n, dtype=torch.uint8, device=dev)
    if rank == 0:
        b = torch.empty(n, dtype=torch.uint8, device=torch.device("cuda", 1))
        for _ in range(2): b.copy_(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): b.copy_(a, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        print("CE P2P copy 64 MiB: %.1f GB/s" % (10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9), flush=True)
        m = 4 << 20
        e0.record()
        for _ in range(10): b[:m].copy_(a[:m], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        print("CE P2P copy 4 MiB: %.1f us each" % (e0.elapsed_time(e1) * 100), flush=True)
dist.barrier()
dist.destroy_process_group()
