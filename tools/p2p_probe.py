"""Diagnostics: peer access, raw peer-copy bandwidth (one process), and NCCL send/recv bandwidth + transport (two ranks)."""
import os, sys, time
import torch
if "RANK" not in os.environ:
    n = torch.cuda.device_count()
    print("devices", n, "peer 0->1", torch.cuda.can_device_access_peer(0, 1) if n > 1 else None)
    if n > 1:
        a = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:0")
        b = torch.empty(64 << 20, dtype=torch.uint8, device="cuda:1")
        for _ in range(3):
            b.copy_(a)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t = time.perf_counter()
        for _ in range(20):
            b.copy_(a)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        print("peer copy 64 MiB: %.1f GB/s" % (20 * 64 / 1024 / (time.perf_counter() - t)))
    sys.exit(0)
import torch.distributed as dist
r, w, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for mb in (1, 21, 63):
    s = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    q = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    peer = r ^ 1
    def once():
        for x in dist.batch_isend_irecv([dist.P2POp(dist.isend, s, peer), dist.P2POp(dist.irecv, q, peer)]):
            x.wait()
    for _ in range(3):
        once()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(20):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 20
    if r == 0:
        print("nccl sendrecv %d MiB each way: %.3f ms, %.1f GB/s per direction" % (mb, dt * 1e3, mb / 1024 / dt))
dist.destroy_process_group()
