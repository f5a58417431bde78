"""Why is H2D from torch-pinned memory 25 GB/s when the link does 55? NUMA placement of the pinned pages."""
import ctypes, os, sys, time
import numpy as np, torch
torch.cuda.init()
def bw(h, d, n):
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(20): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); return n / ((time.perf_counter() - t) / 20) / 1e9
n = 28_000_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print(open("/sys/bus/pci/devices/%s/numa_node" % torch.cuda.get_device_properties(0).pci_bus_id_str if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id_str") else "x").read())
except Exception as e:
    pass
os.system("nvidia-smi topo -m 2>/dev/null | head -12")
h = torch.empty(n, dtype=torch.uint8).pin_memory(); print("torch pin_memory default affinity: %.1f GB/s" % bw(h, d, n))
ncpu = os.cpu_count()
for lo, hi in ((0, ncpu // 4), (ncpu // 4, ncpu // 2), (ncpu // 2, 3 * ncpu // 4), (3 * ncpu // 4, ncpu)):
    try:
        os.sched_setaffinity(0, set(range(lo, hi)))
        h = torch.empty(n, dtype=torch.uint8); h.fill_(1); h = h.pin_memory()
        print(f"pinned after affinity cpus {lo}-{hi-1}: {bw(h, d, n):.1f} GB/s")
    except Exception as e:
        print("affinity", lo, hi, "failed", e)
