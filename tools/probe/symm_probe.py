"""Probe: does torch's symmetric memory (peer pointers over NVLink) work on this box?  torchrun, 2 ranks."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.uint8, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.uint8)
    print(rank, "peer view", peer[:4].tolist(), flush=True)
    hdl.barrier()
except Exception as e:
    print(rank, "symm_mem FAILED:", type(e).__name__, e, flush=True)
print(rank, "can_access_peer", torch.cuda.can_device_access_peer(rank, (rank + 1) % world), flush=True)
dist.barrier()
dist.destroy_process_group()
