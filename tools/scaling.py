"""Multi-GPU bookkeeping of the benchmark: independent streams are sharded one per rank (no data-path collective —
the path is embarrassingly parallel per stream); the only collective is ONE all_gather of a small
per-rank counter struct after the timed window (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist

# counter struct layout (float64)
FRAMES, DEV_MS, E2E_MS, LAUNCHES, PARITY_FAIL, WALL_MS, WALL_E2E_MS, OUTPUTS = range(8)
N_COUNTERS = 8


def stream_seed(rank: int) -> int:
    """Stream i of the multi-stream configs uses clip seed 42 + i."""
    return 42 + rank


def gather_counters(mine: torch.Tensor) -> torch.Tensor:
    """One all_gather of the per-rank counters -> [world, N_COUNTERS] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return mine.view(1, -1).clone()
    world = dist.get_world_size()
    out = torch.empty(world * mine.numel(), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine.contiguous())
    return out.view(world, -1)


def aggregate(allc: torch.Tensor) -> dict:
    """Whole-job numbers: frames summed over ranks, time = max over ranks (never a wall clock of one rank)."""
    a = allc.detach().cpu().double()
    frames = float(a[:, FRAMES].sum())
    t_dev = float(a[:, DEV_MS].max())
    t_e2e = float(a[:, E2E_MS].max())
    return {
        "frames": frames,
        "value_fps": frames / (t_dev * 1e-3) if t_dev > 0 else 0.0,
        "e2e_fps": frames / (t_e2e * 1e-3) if t_e2e > 0 else 0.0,
        "dev_ms": t_dev,
        "e2e_ms": t_e2e,
        "launches": int(a[:, LAUNCHES].sum()),
        "parity_failures": int(a[:, PARITY_FAIL].sum()),
        "wall_ms": float(a[:, WALL_MS].max()),
        "outputs": int(a[:, OUTPUTS].sum()),
    }
