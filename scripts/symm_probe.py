"""Dev tool (torchrun): does torch symmetric memory (peer-mapped buffers + device-side barrier) work on this box?"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1 << 21
buf = symm_mem.empty(n, dtype=torch.uint8, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
print(rank, "rendezvous ok; world", hdl.world_size, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "signal pads", len(hdl.signal_pad_ptrs), flush=True)
buf.fill_(rank + 1)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (n,), torch.uint8)
print(rank, "peer value", int(peer[12345].item()), flush=True)
# timing: barrier, peer copy
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(10): hdl.barrier()
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(100): hdl.barrier()
e1.record(); torch.cuda.synchronize()
tb = e0.elapsed_time(e1) * 10
loc = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(10): loc.copy_(peer)
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(100): loc.copy_(peer)
e1.record(); torch.cuda.synchronize()
tc = e0.elapsed_time(e1) * 10
t0 = time.perf_counter()
for _ in range(100): hdl.barrier()
th = (time.perf_counter() - t0) * 1e4
print(f"rank {rank}: hdl.barrier {tb:.1f} us (host {th:.1f} us)   2 MiB peer copy {tc:.1f} us ({n / tc / 1e3:.0f} GB/s)", flush=True)
dist.destroy_process_group()
