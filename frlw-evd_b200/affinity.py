"""CPU / NUMA placement of a rank next to its GPU.

With one process per GPU the pinned host buffers of the end-to-end path (raw ``.dat`` bytes
in, uint8 tensors out) should live on the NUMA node the GPU hangs off: Linux allocates pages
on the node of the CPU that first touches them, so binding the process to the GPU's CPUs
before anything is allocated is enough.  Without it every rank's buffers can end up on one
socket and the copies of the far GPUs cross the inter-socket link.
"""
from __future__ import annotations

import os


def _cpus_from_sysfs(index: int):
    import pynvml
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
    if len(bus.split(":")[0]) == 8:            # NVML prints an 8-digit PCI domain, sysfs uses 4
        bus = bus[4:]
    with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
        node = int(fh.read())
    if node < 0:
        return None
    with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
        cpus = set()
        for part in fh.read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_device(index: int) -> dict:
    """Bind the calling process to the CPUs local to CUDA device ``index``.  Best effort:
    returns ``{"bound": bool, "cpus": n, "how": ...}`` and never raises (containers may hide
    the topology or forbid the call)."""
    allowed = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            cpus = _cpus_from_sysfs(index)
            how = "sysfs numa_node"
        except Exception:
            cpus, how = None, ""
        if not cpus:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            how = "nvmlDeviceGetCpuAffinity"
        cpus &= allowed
        if not cpus or cpus == allowed:
            return {"bound": False, "cpus": len(allowed), "how": how + " (no narrower set)"}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "cpus": len(cpus), "how": how}
    except Exception as exc:
        return {"bound": False, "cpus": len(allowed), "how": "failed: %r" % (exc,)}
