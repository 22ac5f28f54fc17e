"""Host->device bandwidth per rank under torchrun: every rank alone, then all ranks at once (pinned 1 GiB buffers,
CUDA events).  Explains the flat end-to-end curve of SCALE_r01 (e2e at 4 GPUs ~ e2e at 2): if the concurrent figure
per GPU drops while the solo figure does not, the ranks share a host-side link (PCIe switch uplink / root port /
host memory), and no software on the GPU side can add bandwidth."""
import os, sys, time
import torch, torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 28                                  # 1 GiB of fp32
host = torch.empty(n, dtype=torch.float32).pin_memory()
host.fill_(1.0)
buf = torch.empty(n, dtype=torch.float32, device=dev)

def bw(reps=4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        buf.copy_(host, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return reps * n * 4 / (e0.elapsed_time(e1) / 1e3) / 1e9

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

bw(1)
solo = 0.0
for r in range(world):                       # one rank at a time
    barrier()
    if r == rank:
        solo = bw()
barrier()
together = bw()                              # all ranks at once
barrier()
vals = torch.tensor([solo, together], device=dev, dtype=torch.float64)
allv = [torch.zeros_like(vals) for _ in range(world)]
if world > 1:
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    pci = []
    try:
        import subprocess
        pci = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.split("\n")
    except Exception:
        pass
    print(f"H2D from pinned host memory, 1 GiB copies, world {world}")
    for r, v in enumerate(allv):
        print(f"  rank {r}: alone {float(v[0]):6.1f} GB/s   all {world} ranks at once {float(v[1]):6.1f} GB/s   {pci[r].strip() if r < len(pci) else ''}")
    print(f"  sum: alone (one at a time) {sum(float(v[0]) for v in allv):.1f}   concurrent {sum(float(v[1]) for v in allv):.1f} GB/s")
    if world >= 4:                           # pairs: which ranks share a link?
        print("  pairwise (rank 0 together with rank r):")
for r in range(1, world if world >= 4 else 1):
    barrier()
    v = bw() if rank in (0, r) else 0.0
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    if rank == 0:
        print(f"    0+{r}: rank 0 {float(g[0][0]):6.1f} GB/s, rank {r} {float(g[r][0]):6.1f} GB/s")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
