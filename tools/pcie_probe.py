#!/usr/bin/env python
"""Host<->device copy ceiling of the box: pinned H2D / D2H / bidirectional bandwidth, one process per GPU, all GPUs of a
set copying at the same time (no kernel).  This is the ceiling of bench.py's e2e number: e2e moves every input byte
host->device and every frame byte device->host.

    python tools/pcie_probe.py                       # GPU 0 alone
    python tools/pcie_probe.py --sets 0 0,1,2,3 0,1,2,3,4,5,6,7    # one line per set (same-socket 4, all 8)

Each worker binds to its GPU's NUMA node before allocating (like bench.py / the gateway workers) unless --no-bind, and
reports on which node its pinned pages actually are (/proc/self/numa_maps)."""
import argparse
import json
import multiprocessing as mp
import re
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pages_by_node(addr: int) -> dict:
    """NUMA node -> pages for the mapping that contains `addr` (pinned host memory shows up as a normal VMA)."""
    best = None
    try:
        for line in open("/proc/self/numa_maps"):
            a = int(line.split()[0], 16)
            if a <= addr and (best is None or a > best[0]):
                best = (a, line)
    except OSError:
        return {}
    if not best:
        return {}
    return {int(m.group(1)): int(m.group(2)) for m in re.finditer(r"N(\d+)=(\d+)", best[1])}


def worker(gpu, bind, n_bytes, reps, piece, bar, q):
    from skyplane_b200.numa import bind_to_gpu, gpu_numa_node

    node = bind_to_gpu(gpu) if bind else gpu_numa_node(gpu)
    import torch

    torch.cuda.set_device(gpu)
    h_in = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(2)  # first touch after the affinity is set
    d_a = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            for o in range(0, n_bytes, piece):
                d_a[o:o + piece].copy_(h_in[o:o + piece], non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            for o in range(0, n_bytes, piece):
                h_out[o:o + piece].copy_(d_b[o:o + piece], non_blocking=True)

    def both():
        h2d()
        d2h()

    res = {"gpu": gpu, "gpu_numa_node": node, "pinned_in_pages_by_node": pages_by_node(h_in.data_ptr()),
           "pinned_out_pages_by_node": pages_by_node(h_out.data_ptr())}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("bidir", both)):
        fn()
        torch.cuda.synchronize()
        bar.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        bar.wait()
        res[name + "_gbs"] = n_bytes * reps / dt / 1e9  # bidir: per direction
    q.put(res)


def run_set(gpus, bind, n_bytes, reps, piece):
    ctx = mp.get_context("spawn")
    bar, q = ctx.Barrier(len(gpus)), ctx.Queue()
    ps = [ctx.Process(target=worker, args=(g, bind, n_bytes, reps, piece, bar, q)) for g in gpus]
    for p in ps:
        p.start()
    rows = sorted((q.get(timeout=600) for _ in ps), key=lambda r: r["gpu"])
    for p in ps:
        p.join()
    agg = {k: sum(r[k] for r in rows) for k in ("h2d_gbs", "d2h_gbs", "bidir_gbs")}
    return {"gpus": gpus, "bound_to_numa_node": bind, "bytes_per_gpu": n_bytes, "piece_bytes": piece,
            "aggregate_h2d_gbs": agg["h2d_gbs"], "aggregate_d2h_gbs": agg["d2h_gbs"],
            "aggregate_bidir_each_direction_gbs": agg["bidir_gbs"], "per_gpu": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", nargs="*", default=["0"], help="comma-separated GPU lists, one probe per list")
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--piece-mib", type=int, default=8, help="bytes per cudaMemcpyAsync (8 = one chunk per copy, like sky_submit)")
    ap.add_argument("--no-bind", action="store_true")
    a = ap.parse_args()
    for s in a.sets:
        gpus = [int(x) for x in s.split(",")]
        print(json.dumps(run_set(gpus, not a.no_bind, a.mib << 20, a.reps, a.piece_mib << 20)), flush=True)
