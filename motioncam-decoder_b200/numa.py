"""Host-side placement for the end-to-end feed (hard part 4): eight GPUs pulling compressed frames
over PCIe need ~350 GB/s of host-memory reads, so every rank should run on -- and pin its staging memory on -- the NUMA
node its GPU hangs off.  Pure /sys parsing; no effect on the decode itself."""
import os


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(pci_bus_id):
    """NUMA node of a PCI device ('0000:1b:00.0'), or None when the platform does not say."""
    try:
        with open(f"/sys/bus/pci/devices/{pci_bus_id.lower()}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except (OSError, ValueError):
        return None


def bind_to_gpu_node(pci_bus_id):
    """Restrict the calling process to the CPUs of the GPU's NUMA node (memory then follows first touch).
    Returns a short description of what was done, for the bench's JSON line."""
    node = gpu_numa_node(pci_bus_id)
    if node is None:
        return "numa: unknown node, not bound"
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return f"numa: node {node} has no allowed cpus, not bound"
        os.sched_setaffinity(0, cpus)
        return f"numa: bound to node {node} ({len(cpus)} cpus)"
    except OSError as e:
        return f"numa: bind failed ({e})"
