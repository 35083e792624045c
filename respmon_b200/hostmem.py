"""Host-side placement helpers for the end-to-end path (uploads from pinned memory).

On a multi-socket box every GPU hangs off one NUMA node; a rank whose pinned buffers live on the other socket pulls all
of its uploads across the inter-socket link.  `bind_to_gpu_numa()` pins the calling process to the CPUs of the GPU's node
before the buffers are allocated (first touch then places the pages there).  Best effort: any missing piece of sysfs
leaves the process as it was.
"""
from __future__ import annotations

import os


def _parse_cpulist(text: str) -> set[int]:
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


def gpu_numa_node(device_index: int) -> int | None:
    """NUMA node of a CUDA device from sysfs, or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(device_index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns what was done (for logs)."""
    info = {"device": device_index, "node": None, "cpus": None, "bound": False}
    node = gpu_numa_node(device_index)
    if node is None:
        return info
    info["node"] = node
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
            info["bound"] = True
    except Exception:
        pass
    return info
