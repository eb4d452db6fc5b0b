"""probe: does torch symmetric memory (peer pointers + device barrier) work here?"""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
t = symm_mem.empty(1 << 20, dtype=torch.float64, device="cuda")
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "signal", hex(hdl.signal_pad_ptrs[0]) if hasattr(hdl, "signal_pad_ptrs") else None, flush=True)
t.fill_(rank + 1.0)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float64)
print(rank, "peer value", peer[:2].tolist(), flush=True)
torch.cuda.synchronize()
for ch in range(1):
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    for _ in range(10): hdl.barrier()
    s.record()
    for _ in range(200): hdl.barrier()
    e.record(); torch.cuda.synchronize()
    print(rank, "barrier us", s.elapsed_time(e) / 200 * 1e3, flush=True)
dist.destroy_process_group()
