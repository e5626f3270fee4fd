"""N > 1 path of the benchmark harness on CPU: world_size-2 gloo processes exchange the per-rank counter struct with
one all_gather (NCCL on the GPUs) and agree on the whole-job numbers (frames summed, time = max over ranks)."""
import os
import socket

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools import scaling
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r processed 300 frames in (100 + 20 r) ms on the device, (200 + 50 r) ms end to end
    mine = torch.tensor([300.0, 100.0 + 20 * rank, 200.0 + 50 * rank, 3000.0 + rank, float(rank == 1), 101.0 + 20 * rank,
                         201.0, 300.0], dtype=torch.float64)
    allc = scaling.gather_counters(mine)
    agg = scaling.aggregate(allc)
    assert scaling.stream_seed(rank) == 42 + rank
    out.put((rank, agg, allc.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_counters_gather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    aggs = {r: a for r, a, _ in results}
    assert aggs[0] == aggs[1]  # every rank sees the same whole-job numbers
    a = aggs[0]
    assert a["frames"] == 600.0
    assert abs(a["value_fps"] - 600.0 / 0.120) < 1e-9   # max over ranks of the device time
    assert abs(a["e2e_fps"] - 600.0 / 0.250) < 1e-9
    assert a["launches"] == 6001 and a["parity_failures"] == 1 and a["outputs"] == 600
    rows = results[0][2]
    assert rows[0][1] == 100.0 and rows[1][1] == 120.0  # rank order preserved by the gather


def test_single_process_needs_no_process_group():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools import scaling
    mine = torch.tensor([300.0, 50.0, 80.0, 10.0, 0.0, 51.0, 81.0, 290.0], dtype=torch.float64)
    a = scaling.aggregate(scaling.gather_counters(mine))
    assert a["frames"] == 300.0 and abs(a["value_fps"] - 6000.0) < 1e-9
