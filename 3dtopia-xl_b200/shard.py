"""Data-parallel sharding of batched image-conditioned sampling: one process per GPU, samples split by index.

The reference has no parallelism on this path (inference.py:249 hard-codes cuda:0, inf_bs = 1, :315).  Samples are
independent — no op in the DiT, the sampler or the VAE mixes samples — so sample s goes to rank
s mod G, every rank holds a full weight replica, and NOTHING is exchanged inside the step loop.  Collectives
(NCCL over NVLink on the GPU box, gloo in the CPU tests) appear only at the edges: an optional broadcast of the
noise drawn on rank 0 and the gather of finished latents / voxels to rank 0.

RNG parity: noise is drawn on ONE CPU generator in the reference's order (manual_seed; per image a discarded
randn(1,P,1,4,4,4) then randn(1,P,68): inference.py:250,313,316), so results do not depend on the number of GPUs.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def assigned(num_samples: int, world: int, rank: int) -> List[int]:
    """Indices of the samples rank `rank` owns: s mod world == rank."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, num_samples, world))


def draw_noise(num_samples: int, num_prims: int = 2048, channels: int = 68, seed: int = 42) -> torch.Tensor:
    """x_T for every sample, consuming the CPU generator exactly like the reference's per-image loop."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(num_samples):
        torch.randn(1, num_prims, 1, 4, 4, 4, generator=g)          # drawn and discarded by the reference (only .shape is used)
        out.append(torch.randn(1, num_prims, channels, generator=g))
    return torch.cat(out, 0) if out else torch.empty(0, num_prims, channels)


def gather_to_rank0(local: torch.Tensor, num_samples: int, group=None) -> Optional[torch.Tensor]:
    """Reassemble per-rank results [n_local, ...] (rank r holds samples r, r+G, ...) into sample order on rank 0.
    One all_gather of equally padded blocks; returns None on the other ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (num_samples + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    blocks = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(blocks, pad, group=group)
    if rank != 0:
        return None
    out = torch.empty((num_samples,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = assigned(num_samples, world, r)
        if idx:
            out[idx] = blocks[r][: len(idx)]
    return out


def run_sharded(num_samples: int, per_sample: Callable[[int, torch.Tensor], torch.Tensor], noise: Optional[torch.Tensor] = None,
                seed: int = 42, num_prims: int = 2048, channels: int = 68, device=None, group=None) -> Optional[torch.Tensor]:
    """Run `per_sample(index, x_T[1,P,C]) -> tensor[1,...]` for the samples this rank owns and gather to rank 0.
    `per_sample` is where the DDIM loop (+ decode) of one sample runs; there is no communication inside it."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if noise is None:
        noise = draw_noise(num_samples, num_prims, channels, seed)   # every rank draws the same stream; no broadcast needed
    mine = assigned(num_samples, world, rank)
    outs = [per_sample(s, noise[s:s + 1].to(device) if device is not None else noise[s:s + 1]) for s in mine]
    if outs:
        local = torch.cat(outs, 0)
    else:
        probe = per_sample(0, noise[0:1].to(device) if device is not None else noise[0:1])
        local = probe[:0]
    return gather_to_rank0(local, num_samples, group)


def run_sharded_batched(num_samples: int, per_batch: Callable[[List[int], torch.Tensor], torch.Tensor], batch: int = 1,
                        noise: Optional[torch.Tensor] = None, seed: int = 42, num_prims: int = 2048, channels: int = 68, group=None) -> Optional[torch.Tensor]:
    """Like run_sharded, with the samples a rank owns grouped `batch` at a time into one call
    `per_batch(indices, x_T[len(indices), P, C]) -> tensor[len(indices), ...]` (config #5: 4 samples per GPU share one forward,
    8 sequences under CFG).  The sample -> rank assignment and the noise stream are those of run_sharded, so the result for sample s
    does not depend on `batch` or on the number of ranks beyond the kernels' batch-invariance (tests/test_gpu_dit.py)."""
    if batch < 1:
        raise ValueError("batch must be >= 1")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if noise is None:
        noise = draw_noise(num_samples, num_prims, channels, seed)
    mine = assigned(num_samples, world, rank)
    outs = [per_batch(mine[i:i + batch], noise[mine[i:i + batch]]) for i in range(0, len(mine), batch)]
    if outs:
        local = torch.cat(outs, 0)
    else:                                   # more ranks than samples: this rank still joins the gather with an empty block
        probe = per_batch([0], noise[0:1])
        local = probe[:0]
    return gather_to_rank0(local, num_samples, group)
