# Probe: what does an all-gather / peer copy cost on this box? (run under torchrun)
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (1, 8, 64, 128):
    n = mb * 1024 * 1024 // 8
    full = torch.zeros(n, dtype=torch.float64, device="cuda")
    part = full[rank * (n // world):(rank + 1) * (n // world)]
    for _ in range(3):
        dist.all_gather_into_tensor(full, part)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_gather_into_tensor(full, part)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"allgather total {mb} MiB: {ms*1e3:.1f} us  -> received {(world-1)/world*mb/1024/(ms*1e-3):.1f} GiB/s per rank", flush=True)
if rank == 0 and world >= 2:
    a = torch.zeros(64 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda:0")
    b = torch.zeros_like(a, device="cuda:1")
    print("can_access_peer", torch.cuda.can_device_access_peer(0, 1))
    for _ in range(2): b.copy_(a)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10): b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    print(f"peer copy 64 MiB: {(time.perf_counter()-t)/10*1e3:.3f} ms", flush=True)
dist.barrier()
dist.destroy_process_group()
