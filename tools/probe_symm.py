"""Probe: does torch symmetric memory give peer pointers on this box? (run under torchrun, 2+ GPUs)"""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024, dtype=torch.float32, device=f"cuda:{local}")
    t.fill_(float(rank + 1))
    h = symm.rendezvous(t, dist.group.WORLD.group_name if hasattr(dist.group.WORLD, "group_name") else dist.group.WORLD)
    ptrs = list(h.buffer_ptrs)
    torch.cuda.synchronize(); dist.barrier()
    peer = h.get_buffer((rank + 1) % world, (1024,), torch.float32)
    print(f"rank {rank}: symm ok, ptrs={[hex(p) for p in ptrs]}, peer[0]={float(peer[0])}", flush=True)
except Exception as e:
    print(f"rank {rank}: symmetric memory FAILED: {type(e).__name__}: {e}", flush=True)
print(f"rank {rank}: can_device_access_peer={torch.cuda.can_device_access_peer(local, (local + 1) % world)}", flush=True)
# all_to_all_single sanity
x = torch.arange(world * 4, device="cuda", dtype=torch.float32) + 100 * rank
y = torch.empty_like(x)
dist.all_to_all_single(y, x)
print(f"rank {rank}: a2a {y.tolist()}", flush=True)
dist.destroy_process_group()
