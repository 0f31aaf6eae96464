"""Stream / frame partitioning across the GPUs of one box.

The path has no cross-frame or cross-stream dependency (all statistics are per frame,
color_balance.cpp:396-428), so it shards with no data-path collective: camera stream s runs on
GPU s mod G; a single stream's frames go round-robin with an in-order merge on the host.  The only
exchange is a host-side gather of per-frame detections (a few KB), done with
torch.distributed.gather_object on whatever backend the process group has (gloo on CPU tests,
nccl under the bench).
"""
from typing import Dict, List, Sequence


def streams_for_rank(n_streams: int, rank: int, world: int) -> List[int]:
    """Camera stream s is owned by rank s mod world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return [s for s in range(n_streams) if s % world == rank]


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Single-stream case: frame f is owned by rank f mod world."""
    return streams_for_rank(n_frames, rank, world)


def merge_in_order(per_rank: Sequence[Dict[int, object]]) -> List[object]:
    """Host reorder queue: per-rank {index: result} dicts -> list ordered by index; raises if an
    index is missing or duplicated."""
    merged: Dict[int, object] = {}
    for d in per_rank:
        for k, v in d.items():
            if k in merged:
                raise ValueError("index %d produced by two ranks" % k)
            merged[k] = v
    n = len(merged)
    if sorted(merged) != list(range(n)):
        raise ValueError("indices are not contiguous 0..%d" % (n - 1))
    return [merged[i] for i in range(n)]


def gather_detections(local: Dict[int, object], dst: int = 0):
    """Gathers every rank's {index: detections} on rank `dst` and merges them in order.
    Returns the merged list on `dst`, None elsewhere.  Works without a process group (world 1)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return merge_in_order([local])
    world, rank = dist.get_world_size(), dist.get_rank()
    if dist.get_backend() == "nccl":
        # object collectives on nccl need a device; all_gather_object handles the staging
        bucket = [None] * world
        dist.all_gather_object(bucket, local)
        return merge_in_order(bucket) if rank == dst else None
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst)
    return merge_in_order(bucket) if rank == dst else None


def bind_to_gpu_numa(device_index: int) -> List[int]:
    """Pin the calling process to the CPUs local to GPU `device_index` (its PCIe root's NUMA node), so that
    pinned frame buffers allocated afterwards are first-touched on that node and host<->device copies do
    not cross the socket interconnect.  One module process per camera stream / GPU
    calls this once at start-up.  Returns the CPU list, [] when the topology cannot be read (nothing is
    changed then)."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
        with open(path) as f:
            text = f.read().strip()
        cpus = []
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.extend(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return []
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001  (no sysfs, no permission, old torch: keep the default affinity)
        return []
