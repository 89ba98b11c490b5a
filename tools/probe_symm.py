import os, sys, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024, dtype=torch.uint8, device=f"cuda:{rank}")
    h = symm.rendezvous(t, dist.group.WORLD)
    if rank == 0:
        print("symm ok; attrs:", [a for a in dir(h) if not a.startswith("_")][:40])
        print("buffer_ptrs", h.buffer_ptrs, "signal_pad_ptrs", getattr(h, "signal_pad_ptrs", None))
    t.fill_(rank + 1)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (1024,), torch.uint8)
    print(rank, "peer value", int(peer[0]))
    h.barrier()
except Exception as e:
    print(rank, "symm failed:", repr(e))
print(rank, "can_access_peer", [torch.cuda.can_device_access_peer(rank, j) for j in range(world) if j != rank])
dist.destroy_process_group()
