"""Pattern sharding across the GPUs of one box.

Given P(t) (replicated, tiny) every site pattern's recursion is independent, so rank g owns the contiguous block
[g*N/G, (g+1)*N/G) of the compressed, sorted pattern list (fixed boundaries -> the result is deterministic for a
given G) and only the per-shard scalars are combined: one all-reduce(sum) of the (1 + 2*n_nodes)-vector
(lnL, d1[.], d2[.]) per evaluation, NCCL on GPUs, gloo in the CPU tests.  torch.distributed is plumbing here.
"""
from __future__ import annotations


def shard_range(n_patterns: int, rank: int, world: int):
    """[lo, hi) of the patterns owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return n_patterns * rank // world, n_patterns * (rank + 1) // world


def join_engine(engine, device):
    """Put `engine` (this rank's pattern shard) into the NCCL job of the default process group: rank 0 draws the id
    (bppgpu_comm_unique_id), torch.distributed carries its 128 bytes, every rank calls bppgpu_comm_init.  Afterwards
    bppgpu_eval / bppgpu_eval_device all-reduce (lnL, d1, d2) inside the engine, on the evaluation's stream."""
    import torch
    import torch.distributed as dist
    from . import capi
    rank, world = dist.get_rank(), dist.get_world_size()
    t = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    engine.comm_init(rank, world, t.cpu().numpy().tobytes())


def combine(out_tensor):
    """Sum the per-shard (lnL, d1, d2) vectors in place over the default process group (no-op for world 1).
    Enqueued on the current stream for CUDA tensors: follow bppgpu_eval_device with it, no host sync in between."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out_tensor, op=dist.ReduceOp.SUM)
    return out_tensor
