"""How many independent streams one B200 carries: S host threads, each driving its own StabilizationFilter (own handle,
own CUDA streams) over device-resident 1080p frames with the look-ahead announcement, all on cuda:0.  A single stream is
latency-bound (its tracking chain is a dependent string of small kernels and the host logic sits between frames), so
the aggregate rate grows with S until the remap kernels fill the machine.  Wall-clock over all threads; informational
(the BASELINE configs run one stream per GPU)."""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import livevisionkit_b200 as L
    from tools.synth import Clip

    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--resolution", default="1080p")
    a = ap.parse_args()
    n = a.frames + a.warmup
    clip = Clip(a.resolution, "shake", frames=min(n, 120))
    dev = [torch.from_numpy(clip[i]).cuda() for i in range(len(clip))]  # shared, read-only inputs
    torch.cuda.synchronize()
    for S in a.streams:
        filters = [L.StabilizationFilter(L.StabilizationFilterSettings.obs_homography_preset(), 0) for _ in range(S)]
        outs = [[torch.empty_like(dev[0]) for _ in range(4)] for _ in range(S)]
        start = threading.Barrier(S + 1)
        done = [0] * S

        def work(k):
            try:
                _work(k)
            except BaseException:
                start.abort()  # a failed worker must not leave the others (and the main thread) waiting at the barrier
                raise

        def _work(k):
            s = filters[k].stream
            for i in range(a.warmup):
                s.prefetch(dev[(i + 1) % len(dev)], L.BGR)
                s.submit(dev[i % len(dev)], outs[k][i % 4], L.BGR, i)
            s.sync()
            start.wait(timeout=300)
            for i in range(a.warmup, n):
                s.prefetch(dev[(i + 1) % len(dev)], L.BGR)
                r = s.submit(dev[i % len(dev)], outs[k][i % 4], L.BGR, i)
                done[k] += r.has_output
            s.sync()

        threads = [threading.Thread(target=work, args=(k,)) for k in range(S)]
        for t in threads:
            t.start()
        start.wait(timeout=300)
        t0 = time.perf_counter()
        for t in threads:
            t.join(timeout=300)
        wall = time.perf_counter() - t0
        print(json.dumps({"streams_on_one_gpu": S, "resolution": a.resolution, "frames_per_stream": a.frames,
                          "aggregate_fps": S * a.frames / wall, "per_stream_fps": a.frames / wall, "outputs": sum(done)}))
        for f in filters:
            f.stream.close()


if __name__ == "__main__":
    main()
