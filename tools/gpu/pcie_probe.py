"""Host->device copy bandwidth from pinned memory: default placement vs the GPU's NUMA node (what bounds `e2e`)."""
import os, sys, time, glob
import torch

def gpu_numa_node():
    try:
        bus = torch.cuda.get_device_properties(0).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(0), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(0), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        return int(open(path).read()), path
    except Exception as e:
        return None, repr(e)

def node_cpus(node):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:
        return None

def bw(nbytes, streams=1, reps=5):
    hs = [torch.empty(nbytes // streams, dtype=torch.uint8, pin_memory=True) for _ in range(streams)]
    for h in hs: h.fill_(1)
    ds = [torch.empty(nbytes // streams, dtype=torch.uint8, device="cuda") for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    best = 0
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for h, d, s in zip(hs, ds, ss):
            with torch.cuda.stream(s):
                d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    return best

print("nodes:", sorted(glob.glob("/sys/devices/system/node/node*")), "cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
node, path = gpu_numa_node()
print("gpu numa node:", node, path)
for tag in ("default", "gpu-node"):
    if tag == "gpu-node":
        cpus = node_cpus(node) if node is not None and node >= 0 else None
        if not cpus:
            print("no NUMA info; skipping"); break
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        print("pinning to", len(allowed), "cpus of node", node)
        if not allowed: break
        os.sched_setaffinity(0, allowed)
    for mb in (16, 64, 256, 1024):
        print("%-9s %5d MiB  1 stream %6.1f GB/s   2 streams %6.1f GB/s" % (tag, mb, bw(mb << 20, 1), bw(mb << 20, 2)))
