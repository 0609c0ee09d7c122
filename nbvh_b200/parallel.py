"""Host-side multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  The scene, tree-cut and weights are replicated; rays are independent, so inference shards by pixel
range / sample index with no data-path collective, and training is data parallel over batch slots with one
allreduce of the gradient vector per step."""


def shard_pixel_range(rank, world, n_pixels, align=32):
    """Contiguous pixel range [first, first+count) of `rank`; boundaries are multiples of `align` (the kernel hands
    out 32-pixel chunks), the last rank takes the remainder.  The RNG is keyed by the global pixel id, so any
    partition reproduces the single-GPU frame bit-for-bit."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    chunks = (n_pixels + align - 1) // align
    lo = (chunks * rank) // world
    hi = (chunks * (rank + 1)) // world
    first = lo * align
    end = min(hi * align, n_pixels)
    return first, max(0, end - first)


def shard_batch_slots(rank, world, batch):
    """Training slots [first, first+count) owned permanently by `rank` (slots carry state across steps)."""
    lo = (batch * rank) // world
    hi = (batch * (rank + 1)) // world
    return lo, hi - lo


def allreduce_gradients(grad_tensor, group=None):
    """Sum of the per-rank gradient vectors, in place (NCCL ring/NVLS on the GPU box)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grad_tensor, op=dist.ReduceOp.SUM, group=group)
    return grad_tensor
