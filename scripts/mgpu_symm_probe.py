"""Is NVSwitch multicast available through torch's symmetric memory on this box?  (plumbing probe for multimem.st)"""
import os, sys
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
try:
    import torch.distributed._symmetric_memory as sm
    t = sm.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", lr))
    h = sm.rendezvous(t, dist.group.WORLD)
    mc = getattr(h, "multicast_ptr", None)
    print(f"rank {rank}: buffer_ptrs {[hex(p) for p in h.buffer_ptrs]} multicast_ptr {hex(mc) if mc else mc} signal pads {len(h.signal_pad_ptrs)}", flush=True)
except Exception as e:
    print(f"rank {rank}: symmetric memory failed: {type(e).__name__}: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()
