"""Multi-GPU bookkeeping of the benchmark: independent streams are sharded one per rank (no data-path collective —
the path is embarrassingly parallel per stream); the only collective is ONE all_gather of a small
per-rank counter struct after the timed window (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist

# counter struct layout (float64).  DEV_MS / E2E_MS are the rank's MEDIAN window time; the optional tail carries the
# rank's fastest and slowest window so that the record can tell rank skew from host noise.
FRAMES, DEV_MS, E2E_MS, LAUNCHES, PARITY_FAIL, WALL_MS, WALL_E2E_MS, OUTPUTS = range(8)
N_COUNTERS = 8
DEV_MIN, DEV_MAX, E2E_MIN, E2E_MAX = range(8, 12)
NV12_MS = 12  # optional: median window of the NV12 end-to-end pass


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
    per_rank = None
    if a.shape[1] >= 12:
        frames_r = a[:, FRAMES]
        per_rank = {
            "dev_ms_median": [float(v) for v in a[:, DEV_MS]], "dev_ms_min": [float(v) for v in a[:, DEV_MIN]],
            "dev_ms_max": [float(v) for v in a[:, DEV_MAX]],
            "e2e_ms_median": [float(v) for v in a[:, E2E_MS]], "e2e_ms_min": [float(v) for v in a[:, E2E_MIN]],
            "e2e_ms_max": [float(v) for v in a[:, E2E_MAX]],
            "e2e_fps": [float(f / (t * 1e-3)) if t > 0 else 0.0 for f, t in zip(frames_r, a[:, E2E_MS])],
            "value_fps": [float(f / (t * 1e-3)) if t > 0 else 0.0 for f, t in zip(frames_r, a[:, DEV_MS])],
            "slowest_rank_e2e": int(a[:, E2E_MS].argmax()), "slowest_rank_value": int(a[:, DEV_MS].argmax()),
            # sum of the ranks' own rates: what the job delivers when no rank waits for another (streams are independent)
            "sum_of_rank_e2e_fps": float(sum(f / (t * 1e-3) for f, t in zip(frames_r, a[:, E2E_MS]) if t > 0)),
            "sum_of_rank_value_fps": float(sum(f / (t * 1e-3) for f, t in zip(frames_r, a[:, DEV_MS]) if t > 0)),
        }
    nv12_ms, nv12_fps = None, None
    if a.shape[1] > NV12_MS and float(a[:, NV12_MS].max()) > 0:
        nv12_ms = float(a[:, NV12_MS].max())
        nv12_fps = [float(f / (t * 1e-3)) if t > 0 else 0.0 for f, t in zip(a[:, FRAMES], a[:, NV12_MS])]
    return {
        "nv12_ms": nv12_ms, "nv12_fps_per_rank": nv12_fps,
        "per_rank": per_rank,
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
