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
