"""Frame-pair sharding across ranks: pairs are independent, so one process per GPU takes a
contiguous chunk of the frame stream with |step| frames of overlap at the chunk edge — the same trick the
reference uses between its 512-frame batches (/root/reference/src/denseflow_gpu.cpp:182-188,204-205) — and the
global flow index (base_start, :189) is preserved.  No data-path collective: only a barrier and a MAX over ranks
of the elapsed time / a SUM of the flow counters (reference: total_frames / total_flows, include/dense_flow.h:47-48).
"""
import os


def shard_pairs(n_frames, step, rank, world):
    """Split the M = max(n_frames - |step|, 0) pairs of a stream into `world` contiguous chunks.
    Returns (first_pair, n_pairs, first_frame, n_frames_needed) for `rank`."""
    a = abs(step)
    m = max(n_frames - a, 0)
    base, rem = divmod(m, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    if n == 0:
        return first, 0, min(first, n_frames), 0
    return first, n, first, n + a


def shard_list(items, rank, world):
    """Static i mod G sharding of a video list (unit = one video: keeps the .done marker semantics,
    src/denseflow_gpu.cpp:456-470)."""
    return [it for i, it in enumerate(items) if i % world == rank]


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """torch.distributed over NCCL (GPU box) or gloo (CPU tests); 127.0.0.1 rendezvous from the env."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def all_max(value, device="cpu"):
    """MAX over ranks of a float (the elapsed time of the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(value, device="cpu"):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
