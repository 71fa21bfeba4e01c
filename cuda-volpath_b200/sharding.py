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


def init_nccl_from_torch(renderer, group=None):
    """Bootstrap the library's own NCCL communicator (vp_nccl_init) in a torch.distributed job: rank 0 creates the
    unique id, torch's process group only carries its 128 bytes to the other ranks (plumbing); every later reduce is
    vp_reduce_nccl -- ncclReduce issued by libvolpath_b200.so itself on the stream the caller names."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    raw = renderer.nccl_unique_id() if rank == 0 else bytes(128)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    renderer.nccl_init(world, rank, bytes(t.cpu().tolist()))
    # NCCL connects its rings lazily inside the first collective (hundreds of ms): pay that here, with one float4
    warm = torch.zeros(4, dtype=torch.float32, device="cuda")
    renderer.reduce_nccl(warm.data_ptr(), warm.data_ptr() if rank == 0 else None, 1, root=0, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return rank, world


def split_frames(total_frames, world_size):
    """Strong scaling: `total_frames` sample indices over `world_size` ranks by interleaving; -> frames of each rank."""
    return [frames_for_rank(0, total_frames, r, world_size)[1] for r in range(world_size)]


def init_nccl_via_store(renderer, rank, world, key="volpath_nccl_id"):
    """The same bootstrap without any torch collective: the 128-byte id travels through torch.distributed's TCP store, so
    this may run in a background thread while the main thread builds the volume (ncclCommInitRank + NCCL's lazy ring
    set-up cost seconds at 8 ranks).  Ends with a one-float4 reduce on a private buffer that pays the lazy set-up."""
    import torch.distributed as dist

    store = dist.distributed_c10d._get_default_store()
    if rank == 0:
        store.set(key, renderer.nccl_unique_id())
    raw = bytes(store.get(key))
    renderer.nccl_init(world, rank, raw)
    warm = renderer.L.vp_dev_alloc(16)
    renderer.reduce_nccl(warm, warm if rank == 0 else None, 1, root=0, stream=None)
    renderer.sync()
    renderer.L.vp_dev_free(warm)
    return rank, world
