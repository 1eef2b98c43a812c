"""Frame-parallel sharding across GPUs: the path has no exchange step, so ranks own
contiguous frame ranges, run independently, and the host gathers results by sequence number.
Replaces nn.DataParallel's per-forward scatter / weight broadcast / gather (reference
detect_face/face.py:56).  Works with any torch.distributed backend (nccl on the GPUs, gloo in the CPU
tests); nothing here touches the data path."""
from typing import List, Sequence, Tuple


def shard_range(n_frames: int, rank: int, world: int, batch: int = 1) -> Tuple[int, int]:
    """Contiguous [start, end) of rank `rank`; boundaries fall on multiples of `batch` so every rank
    (except possibly the last) runs full batches."""
    n_batches = -(-n_frames // batch)
    per, extra = divmod(n_batches, world)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    return min(b0 * batch, n_frames), min(b1 * batch, n_frames)


def batches(start: int, end: int, batch: int) -> List[Tuple[int, int]]:
    return [(s, min(s + batch, end)) for s in range(start, end, batch)]


def gather_ordered(local_results: Sequence, start: int, group=None, dst: int = 0):
    """local_results[i] belongs to frame `start + i`.  Returns, on rank `dst`, the list of all results in
    frame order (None elsewhere).  Order is restored from the sequence numbers, not from rank order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(local_results)
    world = dist.get_world_size(group)
    payload = [(start + i, r) for i, r in enumerate(local_results)]
    gathered = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if gathered is None:
        return None
    flat = [item for part in gathered for item in part]
    flat.sort(key=lambda t: t[0])
    seq = [t[0] for t in flat]
    if seq != list(range(len(seq))):
        raise RuntimeError("frame sequence has gaps or duplicates after the gather")
    return [t[1] for t in flat]


def bind_to_gpu_numa_node(device_index: int):
    """Pin the calling process to the CPUs NVML reports as local to CUDA device `device_index` (its NUMA node), BEFORE
    any pinned staging buffer is allocated: with one rank per GPU every rank moves its frames host<->device each step
    (2 x 100 MB per 16-frame 1080p batch), and buffers that were first touched on the other socket make that traffic
    cross the inter-socket link (host PCIe / NUMA is the expected limiter at 8 GPUs).  Host-side plumbing
    only; returns the sorted CPU list now in effect, or None when NVML / the affinity call is unavailable (nothing is
    changed then)."""
    import os
    try:
        import pynvml
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            before = os.sched_getaffinity(0)
            pynvml.nvmlDeviceSetCpuAffinity(h)
            after = os.sched_getaffinity(0)
            if not after or not (after <= before):          # never widen or empty the mask we were started with
                os.sched_setaffinity(0, before)
                return None
            return sorted(after)
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return None
