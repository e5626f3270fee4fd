"""Copy-only ceiling of the end-to-end path: every rank moves frame-sized buffers host->device and device->host
concurrently (pinned memory, two CUDA streams), all ranks at the same time.  The aggregate bytes/s divided by the
bytes one frame moves is the fps no implementation can exceed on this box when every frame crosses PCIe twice.

    python tools/pcie_ceiling.py                      # N = 1
    torchrun --nproc-per-node N tools/pcie_ceiling.py   # N ranks at once
Also imported by bench.py (the `pcie_ceiling` object of the JSON line)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def probe(local, world, frame_bytes, copies=60, dist=None):
    import torch
    dev = torch.device("cuda", local)
    n_buf = 4
    host_in = torch.empty((n_buf, frame_bytes), dtype=torch.uint8).pin_memory()
    host_out = torch.empty((n_buf, frame_bytes), dtype=torch.uint8).pin_memory()
    d_in = torch.empty((n_buf, frame_bytes), dtype=torch.uint8, device=dev)
    d_out = torch.empty((n_buf, frame_bytes), dtype=torch.uint8, device=dev)
    up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h):
        torch.cuda.synchronize()
        if dist is not None and world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(copies):
            if h2d:
                with torch.cuda.stream(up):
                    d_in[i % n_buf].copy_(host_in[i % n_buf], non_blocking=True)
            if d2h:
                with torch.cuda.stream(down):
                    host_out[i % n_buf].copy_(d_out[i % n_buf], non_blocking=True)
        up.synchronize()
        down.synchronize()
        return copies * frame_bytes / (time.perf_counter() - t0) / 1e9

    run(True, True)  # warm-up
    return {"h2d_only_GBps": run(True, False), "d2h_only_GBps": run(False, True), "both_GBps_per_direction": run(True, True)}


def gather(mine, local, world, dist):
    import torch
    t = torch.tensor([mine["h2d_only_GBps"], mine["d2h_only_GBps"], mine["both_GBps_per_direction"]], dtype=torch.float64,
                     device=torch.device("cuda", local))
    if world > 1:
        out = torch.empty(world * 3, dtype=torch.float64, device=t.device)
        dist.all_gather_into_tensor(out, t)
        t = out.view(world, 3)
    else:
        t = t.view(1, 3)
    return t.cpu().tolist()


def summarise(rows, frame_bytes):
    both = [r[2] for r in rows]
    return {"what": "copy-only ceiling measured in this run: every rank copies frame-sized pinned buffers H2D and D2H "
                    "concurrently, all ranks at once (no kernels)",
            "frame_bytes": frame_bytes, "per_rank_both_GBps_per_direction": both,
            "per_rank_h2d_only_GBps": [r[0] for r in rows], "per_rank_d2h_only_GBps": [r[1] for r in rows],
            "aggregate_GBps_per_direction": sum(both),
            "ceiling_fps_whole_job": sum(b * 1e9 / frame_bytes for b in both),
            "ceiling_fps_per_rank": [b * 1e9 / frame_bytes for b in both]}


if __name__ == "__main__":
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fb = 1920 * 1080 * 3
    rows = gather(probe(local, world, fb, dist=dist), local, world, dist)
    if rank == 0:
        print(json.dumps(summarise(rows, fb)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
