"""Sample-index sharding across GPUs.  A path-sample is fully determined by (x, y, frame)
(src/sampler.h:35-43), so rank r of G renders frames {first + r, first + r + G, ...} for all pixels into its own
float4[W*H] sum; the sums are combined by one reduce over NCCL/NVLink.  No data-path collective besides it."""


def frames_for_rank(first_frame, n_frames, rank, world_size):
    """-> (first, count, stride) of the interleaved frame subset of `rank`; the union over ranks is exactly
    first_frame .. first_frame + n_frames - 1, each frame once."""
    if n_frames < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad sharding arguments")
    count = (n_frames - rank + world_size - 1) // world_size if n_frames > rank else 0
    return first_frame + rank, count, world_size


def reduce_accumulators(sum_tensor, dst=0, group=None):
    """Sum the per-rank float4 accumulators onto rank `dst` (torch.distributed: NCCL on GPUs, gloo in CPU
    tests).  In place; returns the tensor.  fp32 addition order differs from a single-GPU run, so the result
    matches it to ~1e-7 relative, not bitwise (the .w scatter counts are integers and match exactly)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(sum_tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return sum_tensor
