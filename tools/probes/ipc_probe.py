"""Feasibility probe (2 GPUs, torchrun): can ranks map each other's device buffers?  Tries torch symmetric memory
and the CUDA-IPC path of torch.multiprocessing.reductions; prints which works and a peer write/read check."""
import os, sys, time, traceback
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
dist.barrier()

def try_symm():
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024, dtype=torch.float32, device=f"cuda:{rank}")
    h = symm.rendezvous(t, dist.group.WORLD.group_name if hasattr(dist.group.WORLD, "group_name") else dist.group.WORLD)
    t.fill_(rank + 1)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (1024,), torch.float32)
    v = peer[0].item()
    h.barrier()
    return f"symm ok: peer value {v}, ptrs {[hex(p) for p in h.buffer_ptrs]}"

def try_ipc():
    from torch.multiprocessing.reductions import reduce_tensor
    t = torch.full((1024,), float(rank + 1), device=f"cuda:{rank}")
    fn, args = reduce_tensor(t)
    objs = [None] * world
    dist.all_gather_object(objs, (fn, args))
    peers = []
    for r, (f, a) in enumerate(objs):
        if r == rank:
            peers.append(t)
        else:
            a = list(a)
            a[6] = rank                      # storage_device: open the handle on THIS rank's device
            peers.append(f(*a))
    torch.cuda.synchronize(); dist.barrier()
    v = peers[(rank + 1) % world][0].item()
    peers[(rank + 1) % world][1] = 100.0 + rank          # peer store
    torch.cuda.synchronize(); dist.barrier()
    got = t[1].item()
    return f"ipc ok: peer value {v}, value stored into me {got}, peer ptr {hex(peers[(rank + 1) % world].data_ptr())} dev {peers[(rank + 1) % world].device}"

which = sys.argv[1:] or ["ipc"]
for name, fn in (("symm", try_symm), ("ipc", try_ipc)):
    if name not in which:
        continue
    print(f"[rank {rank}] trying {name}", flush=True)
    try:
        print(f"[rank {rank}] {fn()}", flush=True)
    except Exception as e:
        print(f"[rank {rank}] {name} FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
        traceback.print_exc()
dist.barrier()
dist.destroy_process_group()
