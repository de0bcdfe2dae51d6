"""Pin a GPU worker process to the CPUs of its GPU's NUMA node (host runtime detail).

With 8 GPUs spread over two sockets, pinned staging buffers that land on the wrong socket make every H2D/D2H
copy cross the inter-socket link.  Called by each worker before it allocates pinned memory, so first-touch
placement follows the affinity.  Best effort: silently does nothing when sysfs / nvidia-smi are unavailable."""
from __future__ import annotations

import os
import subprocess
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
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0].strip()
        dom, bus, rest = out.split(":")
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{bus.lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
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
