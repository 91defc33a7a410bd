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


def gpu_cpu_affinity_nvml(pci_bus_id):
    """The CPUs NVML calls ideal for the GPU (what `nvidia-smi topo -m` prints as CPU Affinity), or None.  Containers often
    hide /sys/bus/pci/devices/*/numa_node (it reads -1) while the driver still knows the topology."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode() if isinstance(pci_bus_id, str) else pci_bus_id)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
            return cpus or None
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return None


_ORIGINAL_AFFINITY = None


def unbind():
    """Give the process back the CPUs it had before bind_to_gpu_node (the CPU baseline uses every host core)."""
    if _ORIGINAL_AFFINITY is not None:
        try:
            os.sched_setaffinity(0, _ORIGINAL_AFFINITY)
        except OSError:
            pass


def bind_to_gpu_node(pci_bus_id):
    """Restrict the calling process to the CPUs of the GPU's NUMA node (memory then follows first touch).
    Returns a short description of what was done, for the bench's JSON line."""
    global _ORIGINAL_AFFINITY
    if _ORIGINAL_AFFINITY is None:
        try:
            _ORIGINAL_AFFINITY = os.sched_getaffinity(0)
        except OSError:
            pass
    node = gpu_numa_node(pci_bus_id)
    if node is None:
        cpus = gpu_cpu_affinity_nvml(pci_bus_id)
        if not cpus:
            return "numa: unknown node, not bound"
        try:
            allowed = os.sched_getaffinity(0)
            use = cpus & allowed
            if not use or use == allowed:
                return f"numa: NVML affinity covers {len(cpus)} cpus ({'all allowed' if use else 'none allowed'}), not bound"
            os.sched_setaffinity(0, use)
            return f"numa: bound to the GPU's NVML cpu affinity ({len(use)} of {len(allowed)} allowed cpus)"
        except OSError as e:
            return f"numa: bind failed ({e})"
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
