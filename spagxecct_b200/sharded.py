"""Multi-GPU driver: one process per GPU (torch.distributed), SNPs sharded by contiguous ranges.

The Step-2 loop is embarrassingly parallel over SNPs: rank 0 owns the Step-1 outputs, ONE
broadcast distributes them (NCCL over NVLink on GPUs, gloo in the CPU tests), every rank tests its own
SNP range with its own Context, and the per-rank result blocks are gathered to rank 0.  No other
collective is used and no reduction crosses GPUs.
"""
import numpy as np


def shard_range(m, world, rank):
    """Contiguous, balanced SNP range [lo, hi) of `rank` (a .bed file is variant-major, so a range of SNPs
    is a range of bytes)."""
    base, rem = divmod(int(m), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_inputs(dist, tensors, src=0):
    """One flattened broadcast of all per-analysis vectors (R, E, Y, covariates, PCs...)."""
    import torch
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.broadcast(flat, src=src)
    out, off = [], 0
    for t in tensors:
        out.append(flat[off:off + t.numel()].reshape(t.shape))
        off += t.numel()
    return out


def gather_results(dist, local_block, m_total, ncol, rank, world, device="cpu"):
    """Gather per-rank row blocks (rows = this rank's SNP range) into the M x ncol matrix on rank 0."""
    import torch
    sizes = [shard_range(m_total, world, r) for r in range(world)]
    maxrows = max(hi - lo for lo, hi in sizes)
    pad = torch.full((maxrows, ncol), float("nan"), dtype=torch.float64, device=device)
    lb = torch.as_tensor(np.ascontiguousarray(local_block), dtype=torch.float64, device=device)
    pad[:lb.shape[0]] = lb
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    out = np.empty((m_total, ncol), dtype=np.float64)
    for r, (lo, hi) in enumerate(sizes):
        out[lo:hi] = bufs[r][:hi - lo].cpu().numpy()
    return out


def run_sharded(dist, rank, world, m_total, ncol, vectors, compute, device="cpu"):
    """vectors: list of tensors valid on rank 0 (same shapes everywhere); compute(lo, hi, vectors) -> (hi-lo, ncol)."""
    vecs = broadcast_inputs(dist, vectors) if world > 1 else vectors
    lo, hi = shard_range(m_total, world, rank)
    block = compute(lo, hi, vecs)
    if world == 1:
        return np.asarray(block)
    return gather_results(dist, block, m_total, ncol, rank, world, device)
