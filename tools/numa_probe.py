"""What the box exposes for GPU <-> CPU locality (NVML affinity masks, NUMA nodes, this process's cpuset).  Tooling."""
import glob
import os

print("cpuset", sorted(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    try:
        print(n, open(n + "/cpulist").read().strip(), open(n + "/meminfo").read().split("\n")[0])
    except Exception as e:
        print(n, e)
try:
    import pynvml
    pynvml.nvmlInit()
    for i in range(pynvml.nvmlDeviceGetCount()):
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, 16)
        cpus = [64 * k + b for k, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        pci = pynvml.nvmlDeviceGetPciInfo(h).busId
        try:
            node = open(f"/sys/bus/pci/devices/{pci.lower()[4:] if len(pci) > 12 else pci.lower()}/numa_node").read().strip()
        except Exception as e:
            node = f"? ({e.__class__.__name__})"
        print("gpu", i, pci, "numa", node, "cpus", (cpus[0], cpus[-1], len(cpus)) if cpus else None)
except Exception as e:
    print("nvml failed:", repr(e))
os.system("nvidia-smi topo -m 2>&1 | head -20")
