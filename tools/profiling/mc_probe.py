"""Does this box give symmetric-memory allocations a multicast (NVLS) mapping?  torchrun --nproc-per-node N ..."""
import os, json, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import torch.distributed._symmetric_memory as symm
t = symm.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", rank))
h = symm.rendezvous(t, dist.group.WORLD)
info = {"rank": rank, "world": world, "multicast_ptr": int(getattr(h, "multicast_ptr", 0) or 0),
        "buffer_ptrs": [hex(int(p)) for p in h.buffer_ptrs][:2], "signal_pad_ptrs": len(getattr(h, "signal_pad_ptrs", []))}
try:
    info["has_multicast_support"] = bool(symm._SymmetricMemory.has_multicast_support(torch._C._distributed_c10d.DeviceType.CUDA if hasattr(torch._C._distributed_c10d, "DeviceType") else "cuda", rank))
except Exception as e:
    info["has_multicast_support"] = f"n/a ({type(e).__name__})"
if rank == 0:
    print(json.dumps(info), flush=True)
dist.barrier()
os._exit(0)
