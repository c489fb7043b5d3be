import os, glob, time, torch, ctypes
p = torch.cuda.get_device_properties(0)
bdf = getattr(p, "pci_bus_id", None)
print("props", p.name, "pci", getattr(p, "pci_domain_id", None), getattr(p, "pci_bus_id", None), getattr(p, "pci_device_id", None))
import subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
print(subprocess.run(["bash", "-c", "lscpu | grep -i numa; nproc; cat /sys/devices/system/node/online"], capture_output=True, text=True).stdout)
q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
print("bus id", q)
path = "/sys/bus/pci/devices/" + q.lower()[4:] + "/numa_node"
print(path, open(path).read().strip() if os.path.exists(path) else "missing")
def bw(node_cpus, label):
    if node_cpus:
        os.sched_setaffinity(0, node_cpus)
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(label, "H2D GB/s", n / dt / 1e9)
    t0 = time.perf_counter()
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(label, "D2H GB/s", n / dt / 1e9)
bw(None, "default affinity")
for node in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    cl = open(node + "/cpulist").read().strip()
    cpus = set()
    for part in cl.split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    bw(cpus, os.path.basename(node))
