"""Concurrent pinned-copy microbenchmark: every rank moves `MB` megabytes H2D, D2H and both at once (two streams) between
its own pinned buffer and its GPU, all ranks at the same time -- the host-memory / PCIe ceiling of the box that the
end-to-end numbers of bench.py run into.  Launch: python tools/pcie_probe.py  or
python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 512
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = MB * (1 << 20) // 8
hin = torch.empty(n, dtype=torch.float64).pin_memory()
hout = torch.empty(n, dtype=torch.float64).pin_memory()
hin.fill_(1.0)
din = torch.empty(n, dtype=torch.float64, device=dev)
dout = torch.ones(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(up, down, reps=4):
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    sync()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return (int(up) + int(down)) * reps * MB / 1024.0 / float(t.item())      # GB/s per rank (slowest rank's clock)


for name, u, d in (("H2D", True, False), ("D2H", False, True), ("H2D+D2H", True, True)):
    run(u, d, 1)
    g = run(u, d)
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"PCIE_PROBE ranks={world} {name}: {g:.1f} GB/s per rank, {g * world:.1f} GB/s aggregate ({MB} MB buffers, pinned)")
try:
    cpus = len(os.sched_getaffinity(0))
    numa = open(f"/sys/bus/pci/devices/{torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), 'pci_bus_id') else ''}/numa_node").read().strip()
except Exception:
    cpus, numa = len(os.sched_getaffinity(0)), "?"
if int(os.environ.get("RANK", "0")) == 0:
    print(f"PCIE_PROBE cpus available to a rank: {cpus}; GPU numa node: {numa}")
if world > 1:
    dist.destroy_process_group()
