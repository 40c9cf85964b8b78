"""Why was e2e at N=1 (plain python) 15 k img/s but 24.6 k per GPU under torchrun (round-1 SCALE)?
Measures the pinned host->device copy rate of one 39.3 MB uint8 batch for buffers allocated (a) as the process starts,
(b) after binding EVERY thread of the process to the GPU-local CPUs (sysfs local_cpulist), and prints the topology."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

n = 32 * 3 * 640 * 640
dev = torch.device("cuda", 0)
torch.cuda.init()
d = torch.empty(n, dtype=torch.uint8, device=dev)


def rate(x, tag):
    for _ in range(3):
        d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        d.copy_(x, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"{tag}: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s", flush=True)


def where():
    cpu = os.sched_getcpu() if hasattr(os, "sched_getcpu") else -1
    return f"cpu {cpu}, affinity {len(os.sched_getaffinity(0))} cpus, torch threads {torch.get_num_threads()}"


pr = torch.cuda.get_device_properties(0)
bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
for f in ("numa_node", "local_cpulist"):
    try:
        print(f, open(f"/sys/bus/pci/devices/{bdf}/{f}").read().strip())
    except OSError as e:
        print(f, "unreadable", e)
print("OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"), "|", where())
x0 = torch.empty(n, dtype=torch.uint8).pin_memory()
x0.random_(0, 255)
rate(x0, "pinned before binding")
print("bind:", bench.bind_to_gpu_numa(0), "|", where())
x1 = torch.empty(n, dtype=torch.uint8).pin_memory()
x1.random_(0, 255)
rate(x1, "pinned after binding all threads")
rate(x0, "the first buffer again")
t0 = time.time()
os.system("numactl -H 2>/dev/null | head -4; nvidia-smi topo -m 2>/dev/null | head -4")
