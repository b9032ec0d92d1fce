"""Does torch symmetric memory work on this box and is there an NVSwitch multicast address?
(run under torchrun with >= 2 ranks; the opt-in multicast Gram path needs both)"""
import os, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    if rank == 0:
        print("symm ok; world", hdl.world_size, "multicast_ptr", hex(hdl.multicast_ptr), "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4])
        print([a for a in dir(hdl) if not a.startswith("_")])
    t.zero_()
    hdl.barrier()
    # write into peer through get_buffer
    peer = hdl.get_buffer((rank + 1) % hdl.world_size, (16,), torch.float32)
    peer.fill_(float(rank + 1))
    hdl.barrier()
    print("rank", rank, "sees", t[:2].tolist())
except Exception as e:
    import traceback; traceback.print_exc()
    print("symm FAILED", repr(e)[:300])
dist.destroy_process_group()
