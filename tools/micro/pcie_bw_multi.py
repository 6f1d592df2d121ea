"""Concurrent host<->device bandwidth of all ranks (one process per GPU), with the process either
left where the launcher put it or bound to the CPUs NVML reports as local to its GPU *before* the
pinned buffers are allocated (first touch places them on that NUMA node).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/micro/pcie_bw_multi.py [--affine]
"""
import json, os, sys, time
import torch, torch.distributed as dist

affine = "--affine" in sys.argv
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
cpus = None
if affine:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
    cpus &= os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
n = 1 << 27  # 512 MiB per buffer
hin = torch.empty(n, dtype=torch.float32).fill_(1.0).pin_memory()
hout = torch.empty(n, dtype=torch.float32).fill_(0.0).pin_memory()
din = torch.empty(n, dtype=torch.float32, device="cuda"); dout = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both(reps):
    for _ in range(reps):
        with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
both(2); torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter(); both(8); torch.cuda.synchronize(); dt = time.perf_counter() - t0
gbs = torch.tensor([8 * n * 4 / dt / 1e9], dtype=torch.float64)
allv = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
if world > 1: dist.all_gather(allv, gbs)
else: allv = [gbs]
if rank == 0:
    per = [round(v.item(), 1) for v in allv]
    print(json.dumps({"affine": affine, "ranks": world, "each_way_gbs_per_rank": per, "each_way_gbs_total": round(sum(per), 1),
                      "cpus_rank0": sorted(cpus) if cpus else None, "host_cpus": os.cpu_count()}))
if world > 1: dist.destroy_process_group()
