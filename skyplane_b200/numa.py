"""Pin a GPU worker process to the CPUs of its GPU's NUMA node (host runtime detail).

With 8 GPUs spread over two sockets, pinned staging buffers that land on the wrong socket make every H2D/D2H
copy cross the inter-socket link.  Called by each worker before it allocates pinned memory, so first-touch
placement follows the affinity.  Best effort: silently does nothing when sysfs or the CUDA library are unavailable."""
from __future__ import annotations

import os
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(device: int) -> Optional[int]:
    """NUMA node of CUDA device `device`.  The bus id comes from the CUDA runtime (sky_device_pci_bus_id), i.e. CUDA's
    device order under CUDA_VISIBLE_DEVICES -- nvidia-smi -i indexes by PCI bus order and ignores that variable."""
    try:
        from skyplane_b200 import native

        dom, bus, rest = native.device_pci_bus_id(device).split(":")
        node = int(open(f"/sys/bus/pci/devices/{dom[-4:]}:{bus}:{rest}/numa_node").read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu(device: int) -> Optional[int]:
    """Restrict this process to the CPUs of `device`'s NUMA node. Returns the node or None."""
    node = gpu_numa_node(device)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        cpus = set(_parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None
