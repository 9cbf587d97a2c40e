"""Multi-GPU plumbing for the sampling path: one process per GPU, independent latent trajectories.

The path shards trivially: every (prompt, seed) trajectory is independent, so
there is NO data-path collective.  The only communication is the one-time weight broadcast that
replaces the reference's DistributedDataParallel constructor broadcast
(tools/inferences/inference_i2vgen_entrance.py:160) and an exit barrier (:250).  Works with the
`nccl` backend on GPUs (NVLink 5 / NVSwitch) and with `gloo` on CPU (tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)   # binds the communicator to this GPU (no barrier() guess)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def shard_items(items, rank, world, mode="partition"):
    """partition: rank r takes items[r::world] (new behaviour: N GPUs = N x throughput);
    replicate: every rank runs the full list, differing only by seed -- the reference's semantics
    (inference_i2vgen_entrance.py:164-170)."""
    items = list(items)
    if mode == "replicate" or world <= 1:
        return items
    if mode != "partition":
        raise ValueError(mode)
    return items[rank::world]


@torch.no_grad()
def broadcast_parameters(module, src=0, bucket_bytes=256 << 20):
    """Broadcast every parameter of `module` from rank `src` in flat buckets (one collective per
    ~256 MB instead of one per tensor).  Returns the number of bytes broadcast."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    params = [p for p in module.parameters()]
    total = 0
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size, total
        if not bucket:
            return
        flat = torch.cat([p.data.reshape(-1) for p in bucket])
        dist.broadcast(flat, src=src)
        off = 0
        for p in bucket:
            n = p.numel()
            p.data.copy_(flat[off:off + n].view_as(p))
            off += n
        total += flat.numel() * flat.element_size()
        bucket, size = [], 0

    for p in params:
        nbytes = p.numel() * p.element_size()
        if bucket and (size + nbytes > bucket_bytes or p.dtype != bucket[0].dtype):
            flush()
        bucket.append(p)
        size += nbytes
    flush()
    if hasattr(module, "invalidate_packed"):
        module.invalidate_packed()
    return total


@torch.no_grad()
def broadcast_packed(module, src=0, bucket_bytes=256 << 20, offload_masters=False):
    """Broadcast the PACKED weight arena of a vgen_b200 module (what its kernels read: fp16 GEMM matrices + fp32
    bias / affine vectors, 2.84 GB for UNetSD_I2VGen instead of the 5.68 GB of fp32 masters) from rank `src` in flat
    buckets over NCCL / NVLink.  Every rank packs its own (possibly meaningless) parameters first, which only fixes
    the layout; the values then come from `src`.  With offload_masters the fp32 masters leave the device afterwards
    (on ranks != src they are stale by construction: only inference through the packed arena is valid there).
    Returns the number of bytes broadcast (0 without an initialised process group)."""
    tensors = [t for _, t in module.packed_tensors()]
    total = 0
    if dist.is_initialized() and dist.get_world_size() > 1:
        by_dtype = {}
        for t in tensors:
            by_dtype.setdefault(t.dtype, []).append(t)
        for dtype, ts in by_dtype.items():
            bucket, size = [], 0

            def flush():
                nonlocal bucket, size, total
                if not bucket:
                    return
                flat = torch.cat([t.reshape(-1) for t in bucket])
                dist.broadcast(flat, src=src)
                off = 0
                for t in bucket:
                    t.copy_(flat[off:off + t.numel()].view_as(t))
                    off += t.numel()
                total += flat.numel() * flat.element_size()
                bucket, size = [], 0

            for t in ts:
                nbytes = t.numel() * t.element_size()
                if bucket and size + nbytes > bucket_bytes:
                    flush()
                bucket.append(t)
                size += nbytes
            flush()
    if offload_masters:
        module.offload_masters()
    return total


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (timings are reported as the max over ranks)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
