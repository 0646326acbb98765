"""Does e2e (host-buffer E_cost calls/s) depend on which CPUs / NUMA node the caller runs on?"""
import os, subprocess, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
p = torch.cuda.get_device_properties(0)
bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
sysfs = Path(f"/sys/bus/pci/devices/{bus}")
print("gpu0", bus, "local_cpulist", (sysfs / "local_cpulist").read_text().strip() if (sysfs / "local_cpulist").exists() else "?",
      "numa_node", (sysfs / "numa_node").read_text().strip() if (sysfs / "numa_node").exists() else "?")
print("cpus allowed:", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), "nodes:", sorted(os.listdir("/sys/devices/system/node"))[:6])
def parse(cl):
    out = []
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part.strip():
            out.append(int(part))
    return out
def run(tag):
    from meanfieldvargp_b200.synthetic import make_free_energy, synthetic_l96
    g = synthetic_l96(500, 16, seed=8192)
    fe = make_free_energy(g)
    x = torch.tensor(g["x0"], dtype=torch.float64).pin_memory().numpy()
    for _ in range(50): fe.E_cost(x)
    w = []
    for _ in range(11):
        t0 = time.perf_counter()
        for _ in range(50): fe.E_cost(x)
        w.append(time.perf_counter() - t0)
    w.sort()
    print(f"{tag}: median {50 / w[5]:.0f} evals/s, best {50 / w[0]:.0f}", flush=True)
if len(sys.argv) > 1:
    os.sched_setaffinity(0, parse(sys.argv[1]))
    run(f"cpus {sys.argv[1]}")
else:
    run("default affinity")
